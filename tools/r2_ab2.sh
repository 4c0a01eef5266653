mkdir -p gpurun_out
run() { n=$1; shift; env "$@" ILF_RUN_LANES=1 timeout 300 python bench.py --quick --steps 20 --warmup 5 --e2e-steps 2 > gpurun_out/b_$n.json 2>gpurun_out/b_$n.err || tail -5 gpurun_out/b_$n.err
  python - "$n" <<'PY'
import json,sys
d=json.load(open('gpurun_out/b_%s.json'%sys.argv[1]))
r=d['roofline']; a=r['all_on']
print(sys.argv[1],'ms', d['ms_per_step'], '| ALL_ON ms', a['ms_per_step'], '| sao', r['per_kernel']['sao']['avg_ms'], a['per_kernel']['sao']['avg_ms'], '| db', r['per_kernel']['deblock']['avg_ms'])
PY
}
for n in 1 2 3 5 10 15 30; do run sao_nseg$n ILF_SAO_NSEG=$n; done
