mkdir -p gpurun_out
run() { n=$1; shift; env "$@" timeout 150 python bench.py --quick --steps 20 --warmup 5 --e2e-steps 2 > gpurun_out/b_$n.json 2>gpurun_out/b_$n.err || tail -5 gpurun_out/b_$n.err
  python - "$n" <<'PY'
import json,sys
d=json.load(open('gpurun_out/b_%s.json'%sys.argv[1]))
r=d['roofline']; a=r['all_on']
print(sys.argv[1],'ms', d['ms_per_step'], 'chain', r['chain']['frac'], '| ALL_ON ms', a['ms_per_step'], 'chain', a['chain']['frac'], '| db', r['per_kernel']['deblock']['avg_ms'])
PY
}
run c740 ILF_DB_CTAS=740
run c800 ILF_DB_CTAS=800
run c860 ILF_DB_CTAS=860
run c925 ILF_DB_CTAS=925
run c1000 ILF_DB_CTAS=1000
