mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
for v in "ILF_RUN_LANES=2 ILF_RUN_LANE_POLICY=1" "ILF_RUN_LANES=3 ILF_RUN_LANE_POLICY=1" "ILF_RUN_LANES=3 ILF_RUN_LANE_POLICY=0" "ILF_RUN_LANES=4 ILF_RUN_LANE_POLICY=1"; do
  n=$(echo $v | tr ' =' '__')
  env $v timeout 300 python bench.py --no-cpu-baseline --e2e-steps 2 > gpurun_out/b_$n.json 2>gpurun_out/b_$n.err || tail -5 gpurun_out/b_$n.err
  python - "$n" <<'PY'
import json,sys
d=json.load(open('gpurun_out/b_%s.json'%sys.argv[1]))
r=d['roofline']; a=r['all_on']
print(sys.argv[1],'value', d['value'], 'ms', d['ms_per_step'], 'chain', r['chain']['frac'], '| ALL_ON ms', a['ms_per_step'], 'chain', a['chain']['frac'], 'launches', d.get('gpu_launches'))
PY
done
