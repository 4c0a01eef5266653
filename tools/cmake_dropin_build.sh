#!/bin/bash
# Builds the reference with the documented CMake option in a scratch copy: copy -> tools/apply_dropin_patch.py ->
# cmake -DILF_B200=ON -> make DecoderApp EncoderApp, then checks that the binaries are linked against libilf_b200.so and reach it
# (without a GPU they stop with the library's "no CUDA device" error; with one, DecoderApp must pass the hash SEI check).
# Needs /root/reference and cmake: build container only.   Usage: tools/cmake_dropin_build.sh [scratch dir]
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
REF=${REF:-/root/reference}
W=${1:-/tmp/vtm_ilf_b200_cmake}
rm -rf "$W"; mkdir -p "$W"
cp -r "$REF" "$W/src"; chmod -R u+w "$W/src"
python "$ROOT/tools/apply_dropin_patch.py" "$W/src" "$ROOT"
mkdir -p "$W/src/build"; cd "$W/src/build"
WNO="-Wno-error=maybe-uninitialized -Wno-error=uninitialized -Wno-error=deprecated-declarations -Wno-error=array-bounds -Wno-error=stringop-overflow -Wno-error=misleading-indentation -Wno-error=nonnull -Wno-error=free-nonheap-object -Wno-error=dangling-pointer -Wno-error=use-after-free -Wno-error=unused-but-set-variable -Wno-error=restrict -Wno-error=stringop-overread -Wno-error=format-overflow -Wno-error=deprecated-copy -Wno-error=mismatched-new-delete -Wno-error=aggressive-loop-optimizations -Wno-error=overloaded-virtual -Wno-error=unused-variable -Wno-error=unused-function"
cmake .. -DCMAKE_BUILD_TYPE=Release -DENABLE_VTM=ON -DILF_B200=ON -DILF_B200_ROOT="$ROOT" -DCMAKE_CXX_FLAGS="-include cstdint -include limits $WNO" > "$W/cmake.log" 2>&1
make -j${JOBS:-8} DecoderApp EncoderApp > "$W/make.log" 2>&1 || { tail -30 "$W/make.log"; exit 1; }
BIN=$(dirname "$(find "$W/src/bin" -name DecoderApp -type f | head -1)")
for app in DecoderApp EncoderApp; do
  ldd "$BIN/$app" | grep -q libilf_b200 || LD_LIBRARY_PATH="$ROOT/vvcsoftware_vtm_b200" ldd "$BIN/$app" | grep -q libilf_b200 || { echo "$app is not linked against libilf_b200.so"; exit 1; }
done
export LD_LIBRARY_PATH="$ROOT/vvcsoftware_vtm_b200:$LD_LIBRARY_PATH"
out=$("$BIN/DecoderApp" -b "$ROOT/tests/golden/streams/intra_416x240.bin" -o /dev/null -d 10 2>&1 || true)
if echo "$out" | grep -q "no CUDA device"; then echo "cmake drop-in build OK: DecoderApp reaches libilf_b200 (no GPU here: $(echo "$out" | grep -o 'ilf_create failed[^)]*)' | head -1))"
elif [ "$(echo "$out" | grep -c '(OK)')" = 8 ]; then echo "cmake drop-in build OK: 8 pictures (OK) through libilf_b200"
else echo "unexpected decoder output:"; echo "$out" | tail -5; exit 1; fi
