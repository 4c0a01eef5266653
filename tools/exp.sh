mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 300 python bench.py --no-cpu-baseline --e2e-steps 2 > gpurun_out/bench_new.json 2>gpurun_out/bench_new.err || tail -5 gpurun_out/bench_new.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_new.json'))
r=d['roofline']
print('value', d['value'], 'ms', d['ms_per_step'], 'chain', r['chain'], 'e2e', d['e2e']['value'])
for k,v in r['per_kernel'].items(): print('  ', k, v)
a=r['all_on']
print('ALL_ON value', a['value'], 'ms', a['ms_per_step'], 'chain', a['chain'])
for k,v in a['per_kernel'].items(): print('  ', k, v)
PY
