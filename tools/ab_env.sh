#!/bin/bash
# A/B environment switches of the library on the device-resident bench (run on the GPU box): tools/ab_env.sh "LABEL VAR=val ..." ...
for spec in "$@"; do
  set -- $spec; label=$1; shift
  env "$@" python bench.py --steps 50 --warmup 3 --e2e-steps 2 --no-cpu-baseline > /tmp/b.json 2>/tmp/b.err || { echo "$label FAILED"; tail -3 /tmp/b.err; continue; }
  python - "$label" <<'PY'
import json,sys
d=json.load(open('/tmp/b.json')); r=d['roofline']
print(f"{sys.argv[1]:12s} value {d['value']:9.0f} ms {d['ms_per_step']:.4f} (instr {r['per_kernel_pass']['ms_per_step']:.4f})  all-on ms {r['all_on']['ms_per_step']:.4f} (instr {r['per_kernel_pass']['all_on_ms_per_step']:.4f}) e2e {d['e2e']['value']:.0f}")
PY
done
