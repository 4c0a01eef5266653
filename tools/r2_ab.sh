#!/bin/bash
# A/B of library variants / env settings on the device-resident bench: tools/r2_ab.sh "label ENV=1 ..." ...
mkdir -p gpurun_out
run() {
  local label=$1; shift
  env "$@" python bench.py --steps 30 --warmup 3 --e2e-steps 1 --quick > gpurun_out/b_$label.json 2>gpurun_out/b_$label.err || { echo "$label FAILED"; tail -3 gpurun_out/b_$label.err; return; }
  python - "$label" <<'PY'
import json,sys
d=json.load(open('gpurun_out/b_%s.json'%sys.argv[1]))
pk=d['roofline']['per_kernel']; ao=d['roofline']['all_on']['per_kernel']
print(f"{sys.argv[1]:16s} value {d['value']:9.0f}  real " + " ".join(f"{k}={v['avg_ms']:.4f}" for k,v in pk.items()) + "  | all-on " + " ".join(f"{k}={v['avg_ms']:.4f}" for k,v in ao.items()), 'chain', d['roofline']['all_on'].get('chain',{}).get('frac'))
PY
}
for spec in "$@"; do run $spec; done
