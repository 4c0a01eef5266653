mkdir -p gpurun_out
run() { n=$1; shift; env "$@" timeout 300 python bench.py --quick --steps 20 --warmup 5 --e2e-steps 2 > gpurun_out/b_$n.json 2>gpurun_out/b_$n.err || tail -5 gpurun_out/b_$n.err
  python - "$n" <<'PY'
import json,sys
d=json.load(open('gpurun_out/b_%s.json'%sys.argv[1]))
r=d['roofline']; a=r['all_on']
print(sys.argv[1],'value', d['value'], 'ms', d['ms_per_step'], 'chain', r['chain']['frac'], '| ALL_ON ms', a['ms_per_step'], 'chain', a['chain']['frac'], '| alf', r['per_kernel']['alf']['avg_ms'], a['per_kernel']['alf']['avg_ms'])
PY
}
run alf32_l1 ILF_B200_LIB=$PWD/variants/libilf_alf32.so ILF_RUN_LANES=1
run alf32_l2 ILF_B200_LIB=$PWD/variants/libilf_alf32.so ILF_RUN_LANES=2
run alf32_l3 ILF_B200_LIB=$PWD/variants/libilf_alf32.so ILF_RUN_LANES=3
