mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_synthetic.py -m gpu -x -q 2>&1 | tail -2
for i in 1 2; do
timeout 150 python bench.py --quick --steps 20 --warmup 5 --e2e-steps 2 > gpurun_out/b_q.json 2>gpurun_out/b_q.err || tail -5 gpurun_out/b_q.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/b_q.json'))
r=d['roofline']; a=r['all_on']
print('ms', d['ms_per_step'], 'chain', r['chain']['frac'], '| ALL_ON ms', a['ms_per_step'], 'chain', a['chain']['frac'], '| db', r['per_kernel']['deblock']['avg_ms'], r['per_kernel']['deblock']['frac'])
PY
done
