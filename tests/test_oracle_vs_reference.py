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
    assert any(any((c["sao_" + k] != c["dbk_" + k]).any() for k in K) for c in caps if "sao_y" in c), "no fixture exercises SAO"
    assert any((c["alf_y"] != c["sao_y"]).any() for c in caps if "alf_y" in c), "no fixture exercises ALF"
    # encoder-side captures (oracle/enc_capture_hook.cpp): three slices that deblocking must NOT cross, beta / tc offsets, chroma QP offsets
    ms = [c for c in caps if c["geom"]["num_slices"] > 1]
    assert ms and all(len(np.unique(c["ctu_slice"])) == 3 for c in ms)
    for c in ms:
        p = np.frombuffer(c["db_params"].tobytes(), np.int32)
        sl = np.frombuffer(c["db_params"].tobytes()[16:16 + 4 * 3], np.int8).reshape(3, 4)
        assert (p[0], p[1]) == (3, -4) and (sl[:, 0] == 2).all() and (sl[:, 1] == -1).all()
        assert any((c["dbk_" + k] != c["pre_" + k]).any() for k in K)


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


ENC_CAPTURE = os.path.join(G.ROOT, "oracle", "_ref", "enc_capture")
ENC_CFG = os.path.join(G.ROOT, "oracle", "_ref", "cfg", "encoder_lowdelay_vtm.cfg")


@pytest.mark.skipif(not (os.path.exists(ENC_CAPTURE) and os.path.exists(ENC_CFG)), reason="oracle/_ref/enc_capture not built (needs /root/reference)")
def test_oracle_matches_fresh_encoder_capture_multi_slice(oracle, tmp_path):
    """The reference ENCODER deblocks pictures its decoder cannot parse: three slices per picture, no filtering across slice
    boundaries, beta / tc offsets, chroma QP offsets.  Every loopFilterPic call of a 2-picture encode: oracle == reference."""
    import sys
    yuv = str(tmp_path / "in.yuv")
    subprocess.run([sys.executable, os.path.join(G.ROOT, "tools", "gen_yuv.py"), "--kind", "small", "-W", "416", "-H", "240", "-n", "2", "--seed", "4321", "-o", yuv], check=True,
                   stdout=subprocess.DEVNULL)
    env = dict(os.environ, ILF_CAPTURE_DIR=str(tmp_path))
    r = subprocess.run([ENC_CAPTURE, "-c", ENC_CFG, "-i", yuv, "-wdt", "416", "-hgt", "240", "-fr", "30", "-f", "2", "--InputBitDepth=10", "--InputChromaFormat=420", "-q", "32",
                        "-b", str(tmp_path / "o.bin"), "-o", str(tmp_path / "rec.yuv"), "--SliceMode=1", "--SliceArgument=3", "--LFCrossSliceBoundaryFlag=0",
                        "--LoopFilterBetaOffset_div2=-2", "--LoopFilterTcOffset_div2=3", "--CbQpOffset=-5", "--CrQpOffset=6"], env=env, capture_output=True, text=True)
    assert r.returncode == 0, (r.stdout + r.stderr)[-2000:]
    files = sorted(glob.glob(str(tmp_path / "pic_*.ilfcap")))
    assert len(files) >= 2
    for f in files:
        c = ilfcap.load(f)
        assert c["geom"]["num_slices"] == 3
        _check(c, oracle)
