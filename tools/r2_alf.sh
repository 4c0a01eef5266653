#!/bin/bash
# Round-2 ALF A/B: (optional) GPU parity tests of the ALF paths, then the device-resident bench with variants.
mkdir -p gpurun_out
if [ -z "$NOTEST" ]; then
python -m pytest tests -m gpu -x -q -k "alf or capture or fullsize or smoke" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
fi
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
run dot X=1
run dot_split ILF_ALF_SPLIT=1
[ -z "$NOGEN" ] && run general_split ILF_ALF_GENERAL=1 ILF_ALF_SPLIT=1
for v in "$@"; do
run ${v} ILF_B200_LIB=$PWD/variants/libilf_$v.so
run ${v}_split ILF_B200_LIB=$PWD/variants/libilf_$v.so ILF_ALF_SPLIT=1
done
