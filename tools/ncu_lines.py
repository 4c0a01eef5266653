#!/usr/bin/env python3
"""Per-source-line executed warp instructions and stall samples of one kernel from an ncu report.
Usage: tools/ncu_lines.py <rep> <kernel regex> [min pct]"""
import csv, io, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
minpct = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
fname, hdr, agg = None, None, {}
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None: continue
    try:
        line = int(r[0]); ie = int(r[hdr.index("Instructions Executed")]); ns = int(r[hdr.index("# Samples")])
    except Exception:
        continue
    if r[2] != "-":   # a SASS row; the CUDA row of the line carries the aggregate
        continue
    k = (fname, line)
    a = agg.setdefault(k, [0, 0, r[1]])
    a[0] += ie; a[1] += ns
tot = sum(a[0] for a in agg.values()); tots = sum(a[1] for a in agg.values())
print(f"total warp instructions {tot}, stall samples {tots}")
for (f, l), (ie, ns, src) in sorted(agg.items(), key=lambda kv: (kv[0][0], kv[0][1])):
    if ie >= tot * minpct / 100 or ns >= tots * minpct / 100:
        print(f"{ie / tot * 100:5.1f}% inst {ns / max(tots,1) * 100:5.1f}% stall  {f}:{l}: {src.strip()[:120]}")
