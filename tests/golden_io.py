"""Loader for tests/golden/captures/*.npz (written by tools/make_golden.py from reference captures)."""
import glob, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tools"), os.path.join(ROOT, "oracle")]
import ilfcap  # noqa: E402

K = ("y", "cb", "cr")
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden", "captures")


def golden_files():
    return sorted(glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load_golden(path):
    """Return the capture as the same dict tools/ilfcap.load gives (planes of every stage reconstructed)."""
    z = np.load(path)
    c = {k: z[k] for k in z.files if not k.startswith("d_")}
    c["geom"] = dict(zip(ilfcap.GEOM_FIELDS, (int(v) for v in z["geom"])))
    prev = "pre"
    for st in ("dbk", "sao", "alf"):
        if f"d_{st}_y" in z.files:
            for k in K:
                c[f"{st}_{k}"] = (c[f"{prev}_{k}"].astype(np.int32) + z[f"d_{st}_{k}"]).astype(np.int16)
            prev = st
    return c


def stage_inputs(c):
    """(name of the stage's input planes) for each stage present."""
    order = [s for s in ("dbk", "sao", "alf") if f"{s}_y" in c]
    prev = "pre"
    out = {}
    for s in order:
        out[s] = prev
        prev = s
    return out
