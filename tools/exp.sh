mkdir -p gpurun_out
for wl in intra_8k_bands ld_1080p_x64; do
timeout 600 python bench.py --workload $wl --steps 20 --no-cpu-baseline --e2e-steps 2 > gpurun_out/bench_$wl.json 2>gpurun_out/bench_$wl.err || tail -8 gpurun_out/bench_$wl.err
python - gpurun_out/bench_$wl.json <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
r=d['roofline']
print(d['config']['workload'][:50], 'value', d['value'], 'ms', d['ms_per_step'], 'chain', r['chain'], 'e2e', d['e2e']['value'], 'all_on', r['all_on']['value'])
for k,v in r['per_kernel'].items(): print('  ', k, v)
PY
done
