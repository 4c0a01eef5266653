mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_alf_stats.py tests/test_gpu_encoder_dropin.py -m gpu -x -q 2>&1 | tail -4
timeout 300 python bench.py --quick --steps 20 --warmup 5 --e2e-steps 2 > gpurun_out/b_q.json 2>gpurun_out/b_q.err || tail -5 gpurun_out/b_q.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/b_q.json'))
print({k:(v['avg_ms'],v['frac']) for k,v in d['next_rows'].items() if 'avg_ms' in v})
PY
