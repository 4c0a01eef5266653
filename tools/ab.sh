#!/bin/bash
# A/B the library variants under variants/ on the device-resident bench (run on the GPU box): tools/ab.sh [name...]
run() {
  local label=$1; shift
  env "$@" python bench.py --steps 30 --warmup 3 --e2e-steps 1 --quick > /tmp/b.json 2>/tmp/b.err || { echo "$label FAILED"; tail -3 /tmp/b.err; return; }
  python - "$label" <<'PY'
import json,sys
d=json.load(open('/tmp/b.json'))
pk=d['roofline']['per_kernel']; ao=d['roofline']['all_on']['per_kernel']
print(f"{sys.argv[1]:24s} value {d['value']:9.0f}  real " + " ".join(f"{k}={v['avg_ms']:.4f}" for k,v in pk.items()) + "  | all-on " + " ".join(f"{k}={v['avg_ms']:.4f}" for k,v in ao.items()))
PY
}
run "base" X=1
for n in "$@"; do run "$n" ILF_B200_LIB=$PWD/variants/libilf_$n.so; done
