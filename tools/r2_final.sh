#!/bin/bash
# Final check of the round on one GPU: the whole GPU suite, smoke(), the full bench line.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err || tail -5 gpurun_out/r02_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err || tail -5 gpurun_out/r02_bench_ref.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench.json'))
r=d['roofline']; a=r['all_on']
print('value', d['value'], 'ms', d['ms_per_step'], 'chain', r['chain']['frac'], '| ALL_ON ms', a['ms_per_step'], 'chain', a['chain']['frac'], '| e2e', d['e2e']['value'], '| cpu', d['cpu_baseline']['value'] if d.get('cpu_baseline') else None, '| launches', d['gpu_launches'], d['clocks'])
print({k:(v['avg_ms'],v['frac']) for k,v in r['per_kernel'].items()}, {k:(v['avg_ms'],v['frac']) for k,v in a['per_kernel'].items()})
print(json.dumps(d.get('dropin'))[:600])
print(open('gpurun_out/r02_bench_ref.json').read()[:600])
PY
