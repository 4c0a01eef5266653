mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'deblock_kernel' -s 3 -c 1 -o gpurun_out/prof_db -f python bench.py --steps 2 --warmup 1 --e2e-steps 1 --no-cpu-baseline > gpurun_out/ncu_db.log 2>&1
tail -1 gpurun_out/ncu_db.log | cut -c1-100
