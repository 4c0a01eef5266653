#!/bin/bash
# One gpurun call: GPU parity tests, smoke, bench, ncu launch list and full captures.  Outputs under gpurun_out/.
mkdir -p gpurun_out
WL=${WL:-ra_4k}
BATCH=${BATCH:-0}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
python bench.py --workload $WL --batch $BATCH > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json
if [ -z "$NO_NCU" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv \
    python bench.py --workload $WL --batch $BATCH --steps 2 --warmup 1 --e2e-steps 1 --quick > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'deblock_kernel|sao_kernel|alf_kernel|sao_stats_kernel' -s 3 -c 3 \
    -o gpurun_out/prof -f python bench.py --workload $WL --batch $BATCH --steps 2 --warmup 1 --e2e-steps 1 --quick > gpurun_out/ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'sao_stats_kernel' -s 1 -c 1 \
    -o gpurun_out/prof_stats -f python bench.py --workload $WL --batch $BATCH --steps 2 --warmup 1 --e2e-steps 1 --quick > gpurun_out/ncu_stats.log 2>&1
fi
tail -3 gpurun_out/pytest_gpu.log
