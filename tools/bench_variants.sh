#!/bin/bash
# Compare library variants / debug modes on the device-resident bench (prints deblock kernel stats only).
run() { # label env...
  local label=$1; shift
  env "$@" python bench.py --steps 20 --warmup 3 --e2e-steps 1 --no-cpu-baseline > /tmp/b.json 2>/tmp/b.err || { echo "$label FAILED"; tail -3 /tmp/b.err; return; }
  python - "$label" <<'PY'
import json,sys
d=json.load(open('/tmp/b.json'))
pk=d['roofline']['per_kernel']
print(sys.argv[1], 'value', d['value'], {k:(v['avg_ms'], v['algo_gbs']) for k,v in pk.items()})
PY
}
run "v3 persistent        " X=1
run "v3 persistent copy   " ILF_DEBUG=1
run "v2 one-shot          " ILF_B200_LIB=$PWD/variants/libilf_v2db.so
run "v2 one-shot copy     " ILF_B200_LIB=$PWD/variants/libilf_v2db.so ILF_DEBUG=1
