#!/usr/bin/env python3
"""SASS instructions of given source lines of one kernel from an ncu report, with executed warp-instruction counts.
Usage: tools/ncu_sass.py <rep> <kernel regex> <first line> <last line>"""
import csv, io, subprocess, sys
rep, kern, l0, l1 = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
cur, hdr = None, None
for r in csv.reader(io.StringIO(raw)):
    if not r: continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < 8: continue
    if r[2] == "-":
        try: cur = int(r[0])
        except ValueError: cur = None
        if cur is not None and l0 <= cur <= l1: print(f"--- {cur}: {r[1].strip()[:110]}   [{r[7]}]")
        continue
    if cur is not None and l0 <= cur <= l1:
        print(f"      {r[7]:>10s}  {r[3].strip()[:100]}")
