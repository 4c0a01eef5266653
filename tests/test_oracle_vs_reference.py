"""Pins the CPU restatement (oracle/ilf_oracle.c) against outputs of the reference itself.

* committed fixtures tests/golden/captures/*.npz: pictures dumped by the unmodified reference decoder
  (oracle/capture_hook.cpp) before deblocking and after each of its three filter stages;
* when oracle/_ref/vtm_capture is present (build container, GPU box): every picture of the small committed
  bitstreams is captured afresh and compared too.
Bit-exact, every stage, every plane."""
import glob
import os
import subprocess

import numpy as np
import pytest

import golden_io as G
import ilfcap

K = G.K


def _check(c, oracle):
    g = c["geom"]
    bd = (g["bd_luma"], g["bd_chroma"], g["ctu_log2"])
    out = oracle.deblock({k: c["pre_" + k] for k in K}, *bd, c["db_params"].tobytes(), c["db_info"], c.get("db_info_c"), None, c["db_mv32"], c["ctu_slice"])
    for k in K:
        assert np.array_equal(out[k], c["dbk_" + k]), f"deblock {k}"
    cur = "dbk"
    if "sao_y" in c:
        s = oracle.sao({k: c["dbk_" + k] for k in K}, *bd, c["sao_ctus"])
        for k in K:
            assert np.array_equal(s[k], c["sao_" + k]), f"sao {k}"
        cur = "sao"
    if "alf_y" in c:
        a = oracle.alf({k: c[f"{cur}_{k}"] for k in K}, *bd, c["alf_params"].tobytes(), c["alf_ctu_enable"])
        for k in K:
            assert np.array_equal(a[k], c["alf_" + k]), f"alf {k}"


@pytest.mark.parametrize("path", G.golden_files(), ids=os.path.basename)
def test_oracle_matches_golden_capture(path, oracle):
    _check(G.load_golden(path), oracle)


def test_golden_set_covers_slice_types_and_tools():
    caps = [G.load_golden(p) for p in G.golden_files()]
    assert {c["geom"]["slice_type"] for c in caps} == {0, 1, 2}            # B, P, I
    assert any(c["geom"]["dual_tree"] for c in caps)
    assert any(any((c["sao_" + k] != c["dbk_" + k]).any() for k in K) for c in caps), "no fixture exercises SAO"
    assert any((c["alf_y"] != c["sao_y"]).any() for c in caps), "no fixture exercises ALF"


CAPTURE = os.path.join(G.ROOT, "oracle", "_ref", "vtm_capture")


@pytest.mark.skipif(not os.path.exists(CAPTURE), reason="oracle/_ref/vtm_capture not built (needs /root/reference)")
@pytest.mark.parametrize("stream", ["intra_416x240", "ra_416x240", "ldp_416x240", "ldb_416x240"])
def test_oracle_matches_fresh_reference_capture(stream, oracle, tmp_path):
    env = dict(os.environ, ILF_CAPTURE_DIR=str(tmp_path))
    r = subprocess.run([CAPTURE, "-b", os.path.join(G.ROOT, "tests", "golden", "streams", stream + ".bin"), "-d", "10", "-o", "/dev/null"],
                       env=env, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "ERROR" not in r.stdout and r.stdout.count("(OK)") > 0      # decoded-picture-hash SEI verified by the reference
    files = sorted(glob.glob(str(tmp_path / "*.ilfcap")))
    assert len(files) == r.stdout.count("(OK)")
    for f in files:
        _check(ilfcap.load(f), oracle)
