mkdir -p gpurun_out
run() { n=$1; shift; env "$@" timeout 150 python bench.py --quick --steps 20 --warmup 5 --e2e-steps 2 > gpurun_out/b_$n.json 2>gpurun_out/b_$n.err || tail -5 gpurun_out/b_$n.err
  python - "$n" <<'PY'
import json,sys
d=json.load(open('gpurun_out/b_%s.json'%sys.argv[1]))
r=d['roofline']; a=r['all_on']
print(sys.argv[1],'ms', d['ms_per_step'], 'chain', r['chain']['frac'], '| ALL_ON ms', a['ms_per_step'], 'chain', a['chain']['frac'])
PY
}
run l2p2 ILF_RUN_LANES=2 ILF_RUN_LANE_POLICY=2
run l3p2 ILF_RUN_LANES=3 ILF_RUN_LANE_POLICY=2
