mkdir -p gpurun_out
N=${N:-2}
nvidia-smi topo -m > gpurun_out/topo_$N.txt 2>&1
for wl in intra_8k_bands ld_1080p_x64 ra_4k; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $N --workload $wl --steps ${STEPS:-50} --no-cpu-baseline --e2e-steps 3 > gpurun_out/mg${N}_$wl.json 2>gpurun_out/mg${N}_$wl.err || tail -12 gpurun_out/mg${N}_$wl.err
python - gpurun_out/mg${N}_$wl.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    r=d['roofline']
    print(d['config']['workload'][:50], 'N', d['n_gpus'], 'value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'all_on', r['all_on']['value'], 'deblock', r['per_kernel']['deblock']['algo_gbs'])
except Exception as e:
    print('no json', e)
PY
done
