#!/usr/bin/env python
"""Experiment: do the chain's kernels gain from mixed residency?  Two contexts (two sets of streams) hold the halves of the
17-picture 4K batch; context B's chain is rotated by one or two stages against A's, so that a memory-bound kernel of one and the
instruction-bound ALF of the other are on the GPU at the same time.  ILF_{DB,SAO,ALF}_SMEM_PAD (environment, read by the library)
cap the CTAs per SM of each kernel so that both kinds fit an SM.  Prints ms per 17-picture step for the single-context run
and for the two-context run."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--rot", type=int, default=1, help="stages context B is ahead of A (0: in phase)")
    ap.add_argument("--all-on", type=int, default=1)
    ap.add_argument("--split", type=int, default=0, help="pictures in context A (0: half)")
    a = ap.parse_args()
    import torch
    import vvcsoftware_vtm_b200 as v
    v.load_library()
    wl = bench.WORKLOADS["ra_4k"]
    w, h = wl["width"], wl["height"]
    side = bench.load_sideinfo("ra_4k")
    if a.all_on:
        side = bench.all_on_sideinfo(side)
    B = len(side)
    planes = bench.synth_planes(w, h, 4, seed=1000)

    def make(slots):
        f = v.InLoopFilter(w, h, 10, 10, 7, device=0, num_slots=len(slots))
        for i, s in enumerate(slots):
            f.upload(i, *planes[s % len(planes)])
            si = side[s]
            f.set_deblock_info(i, si["db_params"].tobytes(), si["db_info"], si.get("db_info_c"), si.get("db_mv16"), None, si["ctu_slice"])
            f.set_sao_params(i, si["sao_ctus"])
            f.set_alf_params(i, si["alf_params"].tobytes(), si["alf_ctu_enable"])
        f.sync()
        return f

    def timed(fn, streams):
        torch.cuda.synchronize()
        ev0 = torch.cuda.Event(enable_timing=True)
        evs = [torch.cuda.Event(enable_timing=True) for _ in streams]
        ev0.record(streams[0])
        for s in streams[1:]:
            s.wait_event(ev0)
        fn()
        for e, s in zip(evs, streams):
            e.record(s)
        torch.cuda.synchronize()
        return max(ev0.elapsed_time(e) for e in evs)

    out = {"steps": a.steps, "rot": a.rot, "all_on": a.all_on,
           "pads": {k: os.environ.get(k) for k in ("ILF_DB_SMEM_PAD", "ILF_SAO_SMEM_PAD", "ILF_ALF_SMEM_PAD")}}
    # one context, the whole batch
    f = make(list(range(B)))
    st = torch.cuda.ExternalStream(f.stream())
    for _ in range(3):
        f.run(0, B, 7)
    f.sync()
    out["single_ms"] = timed(lambda: [f.run(0, B, 7) for _ in range(a.steps)], [st]) / a.steps
    f.close()
    # two contexts
    if a.split:
        # pictures with ALF on anywhere first
        order = sorted(range(B), key=lambda i: -int(side[i]["alf_ctu_enable"].any()))
        sa, sb = order[:a.split], order[a.split:]
    else:
        sa, sb = list(range(0, B, 2)), list(range(1, B, 2))
    fa, fb = make(sa), make(sb)
    sta, stb = torch.cuda.ExternalStream(fa.stream()), torch.cuda.ExternalStream(fb.stream())
    na, nb = len(sa), len(sb)
    head = {0: 0, 1: 1, 2: 3}[a.rot]       # stages B runs ahead of the loop
    tail = 7 & ~head

    def both():
        # B: head, (steps - 1) x (tail, head), tail = `steps` chains, rotated against A's
        if head:
            fb.run(0, nb, head)
        for k in range(a.steps):
            fa.run(0, na, 7)
            if not head:
                fb.run(0, nb, 7)
            elif k < a.steps - 1:
                fb.run(0, nb, tail)
                fb.run(0, nb, head)
        if head:
            fb.run(0, nb, tail)

    for _ in range(2):
        both()
    fa.sync(); fb.sync()
    out["two_ms"] = timed(both, [sta, stb]) / a.steps
    out["na"], out["nb"] = na, nb
    print(json.dumps(out))


if __name__ == "__main__":
    main()
