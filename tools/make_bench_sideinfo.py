#!/usr/bin/env python3
"""Side information of real pictures for bench.py (no sample planes: those are synthetic, see bench.py).

Decodes a committed bitstream with the reference decoder + capture hook (oracle/_ref/vtm_capture, build container
only), keeps the flat side-information arrays the product packer produced for the chosen pictures and stores them
compressed under bench_data/<name>.npz.  The per-picture reference filter times printed by the hook are stored too
(field "ref_us": deblock, SAO, ALF microseconds on this container's CPU, informational).

  tools/make_bench_sideinfo.py ra_4k 0 1 2 3 4 8 12 16
"""
import os, subprocess, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "tools")]
import ilfcap

KEYS = ("db_params", "db_info", "db_info_c", "db_mv32", "ctu_slice", "sao_ctus", "alf_params", "alf_ctu_enable")


def main():
    name, picks = sys.argv[1], [int(a) for a in sys.argv[2:]]
    cap = os.path.join(ROOT, "oracle", "_ref", "vtm_capture")
    out = {}
    with tempfile.TemporaryDirectory() as td:
        env = dict(os.environ, ILF_CAPTURE_DIR=td, ILF_CAPTURE_PLANES="0", ILF_CAPTURE_MAX=str(max(picks) + 1))
        subprocess.run([cap, "-b", os.path.join(ROOT, "tests", "golden", "streams", name + ".bin"), "-d", "10", "-o", "/dev/null"],
                       env=env, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        for j, i in enumerate(picks):
            c = ilfcap.load(os.path.join(td, f"pic_{i:04d}.ilfcap"))
            g = c["geom"]
            out[f"p{j}_geom"] = np.array([g[k] for k in ilfcap.GEOM_FIELDS], np.int32)
            for k in KEYS:
                if k not in c:
                    continue
                v = c[k]
                if k == "db_mv32":
                    assert np.abs(v).max(initial=0) < 32768
                    k, v = "db_mv16", v.astype(np.int16)
                out[f"p{j}_{k}"] = v
    out["num_pictures"] = np.array(len(picks), np.int32)
    path = os.path.join(ROOT, "bench_data", name + ".npz")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path))


if __name__ == "__main__":
    main()
