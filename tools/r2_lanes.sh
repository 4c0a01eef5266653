mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_encoder_dropin.py tests/test_gpu_decoder_dropin.py -m gpu -x -q 2>&1 | tail -5
for i in 1 2 3; do
  timeout 300 python bench.py --quick --steps 20 --warmup 5 --e2e-steps 2 > gpurun_out/b_rep$i.json 2>gpurun_out/b_rep$i.err || tail -5 gpurun_out/b_rep$i.err
  python - "rep$i" <<'PY'
import json,sys
d=json.load(open('gpurun_out/b_%s.json'%sys.argv[1]))
r=d['roofline']; a=r['all_on']
print(sys.argv[1],'value', d['value'], 'ms', d['ms_per_step'], 'chain', r['chain']['frac'], '| ALL_ON ms', a['ms_per_step'], 'chain', a['chain']['frac'], 'launches', d.get('gpu_launches'))
PY
done
