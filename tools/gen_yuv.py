#!/usr/bin/env python3
"""Seeded synthetic 4:2:0 10-bit YUV generator (the inputs of every BASELINE config).

Formulas and seeds are the ones written down in BASELINE.md section 4.  Output is planar
Y, Cb, Cr per frame, little-endian 16-bit, values clipped to [0, 1023].

  tools/gen_yuv.py --kind small  -W 416  -H 240  -n 8  --seed 1234 -o syn_416x240_10b.yuv
  tools/gen_yuv.py --kind tex    -W 1920 -H 1080 -n 32 --seed 2026 -o syn_1920x1080_10b.yuv
"""
import argparse
import numpy as np


def frames_small(W, H, n, seed):
    """416x240 recipe: sinusoid + 32-px checker + N(0,12) noise."""
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:H, 0:W].astype(np.float64)
    cy, cx = np.mgrid[0:H // 2, 0:W // 2].astype(np.float64)
    for f in range(n):
        Y = (512 + 300 * np.sin((x + 3 * f) / 23) * np.cos((y - 2 * f) / 17)
             + 120 * ((x // 32 + y // 32 + f) % 2) + rng.normal(0, 12, (H, W)))
        U = 512 + 200 * np.sin((cx + f) / 11) + rng.normal(0, 6, (H // 2, W // 2))
        V = 512 + 200 * np.cos((cy - f) / 13) + rng.normal(0, 6, (H // 2, W // 2))
        yield Y, U, V


def frames_tex(W, H, n, seed):
    """1080p/4K/8K recipe: low-passed noise texture translating (2,1) px per frame."""
    rng = np.random.default_rng(seed)
    th, tw = H + 64, W + 64
    g = rng.normal(0, 1, (th, tw))
    ky = np.fft.fftfreq(th)[:, None]
    kx = np.fft.rfftfreq(tw)[None, :]
    k = np.sqrt(ky * ky + kx * kx)
    T = np.fft.irfft2(np.fft.rfft2(g) / (1 + (60 * k) ** 2), s=(th, tw))
    T /= T.std()
    y, x = np.mgrid[0:H, 0:W].astype(np.float64)
    for f in range(n):
        oy, ox = f % 64, (2 * f) % 64
        Tw = T[oy:oy + H, ox:ox + W]
        Y = (512 + 220 * Tw + 60 * np.sin((x - 4 * f) / 37) + 40 * ((x // 64 + y // 64) % 2)
             + rng.normal(0, 6, (H, W)))
        U = 512 + 120 * Tw[::2, ::2] + rng.normal(0, 3, (H // 2, W // 2))
        V = 512 - 100 * Tw[::2, ::2] + rng.normal(0, 3, (H // 2, W // 2))
        yield Y, U, V


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--kind", choices=["small", "tex"], default="tex")
    ap.add_argument("-W", type=int, required=True)
    ap.add_argument("-H", type=int, required=True)
    ap.add_argument("-n", type=int, required=True)
    ap.add_argument("--seed", type=int, required=True)
    ap.add_argument("-o", required=True)
    a = ap.parse_args()
    gen = frames_small if a.kind == "small" else frames_tex
    with open(a.o, "wb") as fh:
        for planes in gen(a.W, a.H, a.n, a.seed):
            for p in planes:
                np.clip(np.rint(p), 0, 1023).astype("<u2").tofile(fh)


if __name__ == "__main__":
    main()
