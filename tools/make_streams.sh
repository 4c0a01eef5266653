#!/bin/bash
# Regenerates the bitstreams committed under tests/golden/streams/ from seeded synthetic YUV with the
# UNMODIFIED reference encoder (oracle/_ref/EncoderApp, built by oracle/ref.mk).  Needs
# /root/reference (cfg files) -- runs in the build container only, never on the GPU box.
# Usage: tools/make_streams.sh <name> ...   (names below; no args = the quick ones)
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
ENC=$ROOT/oracle/_ref/EncoderApp
CFG=${REF:-/root/reference}/cfg
W=$ROOT/build/streams; mkdir -p $W $ROOT/tests/golden/streams
PY=${PYTHON:-python}

enc() { # name cfg W H frames kind seed qp extra...
  local name=$1 cfg=$2 w=$3 h=$4 n=$5 kind=$6 seed=$7 qp=$8; shift 8
  local yuv=$W/syn_${w}x${h}_s${seed}_n${n}.yuv
  [ -f $yuv ] || $PY $ROOT/tools/gen_yuv.py --kind $kind -W $w -H $h -n $n --seed $seed -o $yuv
  $ENC -c $CFG/$cfg -i $yuv -wdt $w -hgt $h -fr 30 -f $n --InputBitDepth=10 --InputChromaFormat=420 \
       --SEIDecodedPictureHash=1 -q $qp -b $W/$name.bin -o $W/${name}_rec.yuv "$@" > $W/$name.enc.log 2>&1
  cp $W/$name.bin $ROOT/tests/golden/streams/$name.bin
  echo "$name done: $(stat -c %s $W/$name.bin) bytes"
}

for s in "${@:-intra_416x240 ra_416x240 ldp_416x240 ldb_416x240}"; do
case $s in
  intra_416x240) enc $s encoder_intra_vtm.cfg        416  240  8 small 1234 37 --TemporalSubsampleRatio=1 ;;
  ra_416x240)    enc $s encoder_randomaccess_vtm.cfg 416  240 17 small 1235 32 ;;
  ldp_416x240)   enc $s encoder_lowdelay_P_vtm.cfg   416  240  6 small 1236 32 ;;
  ldb_416x240)   enc $s encoder_lowdelay_vtm.cfg     416  240  6 small 1237 30 ;;
  ra_1080p)      enc $s encoder_randomaccess_vtm.cfg 1920 1080 32 tex 2026 37 ;;
  ra_4k)         enc $s encoder_randomaccess_vtm.cfg 3840 2160 17 tex 4000 37 ;;
  intra_8k)      enc $s encoder_intra_vtm.cfg        7680 4320  1 tex 8000 37 --TemporalSubsampleRatio=1 ;;
  ld_1080p_s*)   seed=${s#ld_1080p_s}; enc $s encoder_lowdelay_vtm.cfg 1920 1080 9 tex $seed 37 ;;
  *) echo "unknown stream $s"; exit 1 ;;
esac
done
