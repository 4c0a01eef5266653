#!/usr/bin/env python3
"""Compare the CPU restatement (oracle/liboracle.so) with a directory of reference captures, stage by stage.
Usage: tools/check_capture.py <capture dir> ...   (exit 1 on any mismatch)"""
import glob, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tools"), os.path.join(ROOT, "oracle")]
import ilfcap
import ilf_oracle as O

K = ("y", "cb", "cr")


def check(path, verbose=True):
    c = ilfcap.load(path); g = c["geom"]
    bd = (g["bd_luma"], g["bd_chroma"], g["ctu_log2"])
    bad = 0
    out = O.deblock({k: c["pre_" + k] for k in K}, *bd, c["db_params"].tobytes(), c["db_info"], c.get("db_info_c"), None, c["db_mv32"], c["ctu_slice"])
    d = [int((out[k] != c["dbk_" + k]).sum()) for k in K]
    cur = "dbk_"
    msg = f"deblock diff={d} changed={int((c['pre_y'] != c['dbk_y']).sum())}"
    bad += sum(d)
    if "sao_y" in c:
        s = O.sao({k: c["dbk_" + k] for k in K}, *bd, c["sao_ctus"])
        d = [int((s[k] != c["sao_" + k]).sum()) for k in K]
        msg += f" | sao diff={d} changed={sum(int((c['sao_'+k] != c['dbk_'+k]).sum()) for k in K)}"
        bad += sum(d); cur = "sao_"
    if "alf_y" in c:
        a = O.alf({k: c[cur + k] for k in K}, *bd, c["alf_params"].tobytes(), c["alf_ctu_enable"])
        d = [int((a[k] != c["alf_" + k]).sum()) for k in K]
        msg += f" | alf diff={d} changed={sum(int((c['alf_'+k] != c[cur+k]).sum()) for k in K)}"
        bad += sum(d)
    if verbose:
        print(f"{path}: poc={g['poc']} type={'BPI'[g['slice_type']]} dual={g['dual_tree']} {msg}")
    return bad


if __name__ == "__main__":
    total = 0
    for d in sys.argv[1:]:
        for f in sorted(glob.glob(os.path.join(d, "*.ilfcap"))):
            total += check(f)
    print("MISMATCHES:", total)
    sys.exit(1 if total else 0)
