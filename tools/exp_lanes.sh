#!/bin/bash
# Mixed-residency experiment matrix (tools/exp_lanes.py); results in gpurun_out/lanes.txt
mkdir -p gpurun_out
out=gpurun_out/lanes.txt; : > $out
run() { echo "## $*" >> $out; env "$@" 2>>gpurun_out/lanes.err | tail -1 >> $out; }
P="python tools/exp_lanes.py --steps 20"
run X=0 $P --rot 0
run X=0 $P --rot 1
run X=0 $P --rot 2
run ILF_DB_SMEM_PAD=12000 $P --rot 1
run ILF_DB_SMEM_PAD=12000 $P --rot 2
run ILF_DB_SMEM_PAD=12000 ILF_SAO_SMEM_PAD=26000 $P --rot 1
run ILF_DB_SMEM_PAD=12000 ILF_SAO_SMEM_PAD=26000 $P --rot 2
run ILF_DB_SMEM_PAD=12000 ILF_SAO_SMEM_PAD=26000 ILF_ALF_SMEM_PAD=28000 $P --rot 2
run X=0 $P --all-on 0 --split 4 --rot 0
run X=0 $P --all-on 0 --split 5 --rot 0
run ILF_DB_SMEM_PAD=12000 $P --all-on 0 --split 4 --rot 0
cat $out
