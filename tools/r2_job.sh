mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
run() { n=$1; shift; env "$@" timeout 150 python bench.py --quick --steps 20 --warmup 5 --e2e-steps 2 > gpurun_out/b_$n.json 2>gpurun_out/b_$n.err || tail -5 gpurun_out/b_$n.err
  python - "$n" <<'PY'
import json,sys
d=json.load(open('gpurun_out/b_%s.json'%sys.argv[1]))
r=d['roofline']; a=r['all_on']
print(sys.argv[1],'ms', d['ms_per_step'], 'chain', r['chain']['frac'], '| ALL_ON ms', a['ms_per_step'], 'chain', a['chain']['frac'], '| db', r['per_kernel']['deblock']['avg_ms'], r['per_kernel']['deblock']['frac'])
PY
}
for q in 4 8 12 24 48; do run q$q ILF_DB_QUOTA=$q; done
