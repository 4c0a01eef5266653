#!/bin/bash
# Build a library variant with extra nvcc flags into variants/libilf_<name>.so (git-ignored; travels with gpurun) for
# A/B runs with ILF_B200_LIB.  Usage: tools/mkvariant.sh <name> [nvcc flags...]
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
obj=/tmp/ilf_variant_$name; mkdir -p $obj $root/variants
cd $root/vvcsoftware_vtm_b200/csrc
for f in ilf_api ilf_deblock ilf_sao ilf_alf ilf_sao_stats ilf_hash ilf_alf_stats; do
  nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -I../../include "$@" -c $f.cu -o $obj/$f.o &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $root/variants/libilf_$name.so $obj/*.o
echo built variants/libilf_$name.so
