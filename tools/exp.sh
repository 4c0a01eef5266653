mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for dbg in 0 1; do
  ILF_DEBUG=$dbg python bench.py --steps 30 --no-cpu-baseline --e2e-steps 1 > /tmp/b.json 2>/tmp/b.err || { echo FAILED; tail -3 /tmp/b.err; }
  python - "dbg=$dbg" <<'PY'
import json,sys
d=json.load(open('/tmp/b.json'))
pk=d['roofline']['per_kernel']
print(sys.argv[1], 'value', d['value'], 'deblock', pk['deblock']['algo_gbs'])
PY
done
