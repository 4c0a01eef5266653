#!/bin/bash
# Per-picture filter time inside the reference decoder: stock CPU filters (vtm_capture) vs the GPU drop-in (run on the GPU box).
for s in ra_1080p ra_4k; do
  b=tests/golden/streams/$s.bin
  ILF_TIMING=1 oracle/_ref/vtm_capture -b $b -o /dev/null -d 10 2> /tmp/cpu.txt > /dev/null
  ILF_TIMING=1 oracle/_ref/DecoderApp_ilf_b200 -b $b -o /dev/null -d 10 2> /tmp/gpu.txt > /dev/null
  python - $s <<'PY'
import re,sys
def tot(p):
    us=[tuple(int(v) for v in m) for m in re.findall(r"deblock_us=(\d+) sao_us=(\d+) alf_us=(\d+)", open(p).read())]
    us=us[1:]   # the first picture carries the one-time CUDA context and ilf_create (about 1.7 s)
    n=len(us); return n, [sum(u[k] for u in us)/n/1e3 for k in range(3)]
n,c=tot('/tmp/cpu.txt'); m,g=tot('/tmp/gpu.txt')
print(f"{sys.argv[1]}: {n} pictures after the first; CPU filters ms/picture deblock {c[0]:.2f} sao {c[1]:.2f} alf {c[2]:.2f} total {sum(c):.2f} | GPU drop-in (pack + upload + kernels + download) deblock {g[0]:.2f} sao {g[1]:.2f} alf {g[2]:.2f} total {sum(g):.2f}")
PY
done
