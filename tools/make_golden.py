#!/usr/bin/env python3
"""Regenerate tests/golden/captures/*.npz from the committed bitstreams with the REFERENCE decoder.

Needs oracle/_ref/vtm_capture (built from /root/reference by `make -C oracle ref`); runs in the build container.
Each fixture is one picture: geometry, the flat side information the product packer produced, the picture before
deblocking and -- delta-coded against the previous stage so the file stays small -- the reference's output after
deblocking, SAO and ALF.  tests/golden_io.py loads them back.
"""
import os, subprocess, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "tools")]
import ilfcap

PICKS = {  # stream -> decode-order picture indices
    "intra_416x240": [0],
    "ra_416x240": [0, 1, 5, 9],
    "ldp_416x240": [1, 4],
    "ldb_416x240": [4],
}
# Encoder-side captures (oracle/enc_capture_hook.cpp): pictures only the reference ENCODER filters -- the decoder cannot parse
# multi-slice streams.  name -> (cfg, frames, seed, qp, extra encoder options, loopFilterPic call indices to keep)
ENC_RUNS = {
    "ms3_416x240": ("encoder_lowdelay_vtm.cfg", 4, 1236, 30,
                    ["--SliceMode=1", "--SliceArgument=3", "--LFCrossSliceBoundaryFlag=0", "--LoopFilterBetaOffset_div2=2", "--LoopFilterTcOffset_div2=-1",
                     "--CbQpOffset=3", "--CrQpOffset=-4"], [0, 1]),
}
K = ("y", "cb", "cr")


def to_fixture(c2):
    g = c2["geom"]
    d = {"geom": np.array([g[k] for k in ilfcap.GEOM_FIELDS], np.int32)}
    prev = "pre"
    for st in ("dbk", "sao", "alf"):  # delta chain pre -> dbk -> sao -> alf
        if st + "_y" in c2:
            for k in K:
                d[f"d_{st}_{k}"] = (c2[f"{st}_{k}"].astype(np.int32) - c2[f"{prev}_{k}"].astype(np.int32)).astype(np.int16)
            prev = st
    for k, v in c2.items():
        if k == "geom" or (k[:4] in ("dbk_", "sao_", "alf_") and k[4:] in K):
            continue
        d[k] = v
    return d


def encoder_fixtures(out_dir):
    enc = os.path.join(ROOT, "oracle", "_ref", "enc_capture")
    cfg_dir = os.path.join(os.environ.get("REF", "/root/reference"), "cfg")
    for name, (cfg, frames, seed, qp, extra, picks) in ENC_RUNS.items():
        with tempfile.TemporaryDirectory() as td:
            yuv = os.path.join(td, "in.yuv")
            subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_yuv.py"), "--kind", "small", "-W", "416", "-H", "240", "-n", str(frames), "--seed", str(seed), "-o", yuv],
                           check=True, stdout=subprocess.DEVNULL)
            env = dict(os.environ, ILF_CAPTURE_DIR=td, ILF_CAPTURE_MAX=str(max(picks) + 1))
            subprocess.run([enc, "-c", os.path.join(cfg_dir, cfg), "-i", yuv, "-wdt", "416", "-hgt", "240", "-fr", "30", "-f", str(frames), "--InputBitDepth=10",
                            "--InputChromaFormat=420", "-q", str(qp), "-b", os.path.join(td, "o.bin"), "-o", os.path.join(td, "rec.yuv")] + extra,
                           env=env, check=True, stdout=subprocess.DEVNULL)
            for i in picks:
                path = os.path.join(out_dir, f"{name}_enc_{i:02d}.npz")
                np.savez_compressed(path, **to_fixture(ilfcap.load(os.path.join(td, f"pic_{i:04d}.ilfcap"))))
                print(path, os.path.getsize(path))


def main():
    out_dir = os.path.join(ROOT, "tests", "golden", "captures")
    os.makedirs(out_dir, exist_ok=True)
    cap = os.path.join(ROOT, "oracle", "_ref", "vtm_capture")
    for stream, picks in PICKS.items():
        with tempfile.TemporaryDirectory() as td:
            env = dict(os.environ, ILF_CAPTURE_DIR=td, ILF_CAPTURE_MAX=str(max(picks) + 1))
            subprocess.run([cap, "-b", os.path.join(ROOT, "tests", "golden", "streams", stream + ".bin"), "-d", "10", "-o", "/dev/null"],
                           env=env, check=True, stdout=subprocess.DEVNULL)
            for i in picks:
                d = to_fixture(ilfcap.load(os.path.join(td, f"pic_{i:04d}.ilfcap")))
                path = os.path.join(out_dir, f"{stream}_{i:02d}.npz")
                np.savez_compressed(path, **d)
                print(path, os.path.getsize(path))
    if os.path.exists(os.path.join(ROOT, "oracle", "_ref", "enc_capture")):
        encoder_fixtures(out_dir)


if __name__ == "__main__":
    main()
