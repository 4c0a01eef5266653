"""Drop-in gate, encoder side: the reference EncoderApp linked with the host shim, so that EncGOP::compressGOP's
LoopFilter::loopFilterPic (EncGOP.cpp:2122), the statistics pass of EncSampleAdaptiveOffset::SAOProcess (:2135,
EncSampleAdaptiveOffset.cpp:227) and the classification + covariance passes of EncAdaptiveLoopFilter::ALFProcess
(EncAdaptiveLoopFilter.cpp:257-260) and the filter application at the end of its search (:433-462) run on libilf_b200.so, must write the same bitstream and the same reconstruction, byte for byte, as the stock encoder: every
decision the encoder takes after deblocking (SAO statistics, ALF covariances, reference pictures of later frames)
depends on every deblocked sample."""
import hashlib
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")
SHIM_ENC, STOCK_ENC, STOCK_DEC = (os.path.join(REF, n) for n in ("EncoderApp_ilf_b200", "EncoderApp", "DecoderApp"))
pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not (os.path.exists(SHIM_ENC) and os.path.exists(STOCK_ENC) and os.path.exists(os.path.join(REF, "cfg", "encoder_intra_vtm.cfg"))),
                                                  reason="oracle/_ref encoders not built (needs /root/reference at build time)")]


def _md5(path):
    with open(path, "rb") as f:
        return hashlib.md5(f.read()).hexdigest()


@pytest.mark.parametrize("cfg,frames,qp,extra", [("encoder_intra_vtm.cfg", 2, 37, ["--TemporalSubsampleRatio=1"]), ("encoder_lowdelay_vtm.cfg", 3, 35, [])])
def test_encoder_with_gpu_deblocking_writes_the_same_stream(cfg, frames, qp, extra, tmp_path):
    yuv = str(tmp_path / "in.yuv")
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_yuv.py"), "--kind", "small", "-W", "416", "-H", "240", "-n", str(frames), "--seed", "77", "-o", yuv], check=True)
    outs = {}
    for tag, enc in (("gpu", SHIM_ENC), ("cpu", STOCK_ENC)):
        bit, rec = str(tmp_path / f"{tag}.bin"), str(tmp_path / f"{tag}_rec.yuv")
        r = subprocess.run([enc, "-c", os.path.join(REF, "cfg", cfg), "-i", yuv, "-wdt", "416", "-hgt", "240", "-fr", "30", "-f", str(frames), "--InputBitDepth=10",
                            "--InputChromaFormat=420", "--SEIDecodedPictureHash=1", "-q", str(qp), "-b", bit, "-o", rec] + extra,
                           capture_output=True, text=True, env=dict(os.environ, ILF_TIMING="1"))
        assert r.returncode == 0, (r.stdout + r.stderr)[-3000:]
        if tag == "gpu":
            assert r.stderr.count("deblock_us=") >= frames, r.stderr[-2000:]      # every picture's deblocking went through the CUDA library ...
            assert r.stderr.count("sao_stats_us=") >= frames, r.stderr[-2000:]    # ... and so did the statistics pass of its SAO search
            assert r.stderr.count("alf_stats_us=") >= frames, r.stderr[-2000:]    # ... and the classification + covariances of its ALF search
            assert r.stderr.count("alf_apply_us=") >= 1, r.stderr[-2000:]         # ... and the ALF filters of the pictures whose search switched ALF on
        outs[tag] = (_md5(bit), _md5(rec))
    assert outs["gpu"] == outs["cpu"]
    # and the stock decoder accepts the GPU-encoded stream with every picture hash (OK)
    r = subprocess.run([STOCK_DEC, "-b", str(tmp_path / "gpu.bin"), "-o", str(tmp_path / "dec.yuv"), "-d", "10"], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.count("(OK)") == frames, r.stdout[-2000:]
    assert _md5(str(tmp_path / "dec.yuv")) == outs["cpu"][1]
