"""Drop-in gate: the reference DecoderApp, linked with the host shim (vvcsoftware_vtm_b200/shim/ilf_shim.cpp) so that
LoopFilter::loopFilterPic / SampleAdaptiveOffset::SAOProcess / AdaptiveLoopFilter::ALFProcess run on libilf_b200.so,
decodes the committed bitstreams; the decoder's own decoded-picture-hash SEI check (DecLib.cpp:579-588,
PicYuvMD5.cpp:225-290) must say (OK) for every picture and the written YUV must equal the stock decoder's."""
import hashlib
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM_DEC = os.path.join(ROOT, "oracle", "_ref", "DecoderApp_ilf_b200")
STOCK_DEC = os.path.join(ROOT, "oracle", "_ref", "DecoderApp")
pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not os.path.exists(SHIM_DEC), reason="oracle/_ref/DecoderApp_ilf_b200 not built (needs /root/reference at build time)")]

# every committed stream: BASELINE configs 1 (416x240), 2 (1080p RA), 3 (4K RA), 4 (the 8K intra picture) and 5 (1080p low delay)
STREAMS = {"intra_416x240": 8, "ra_416x240": 17, "ldp_416x240": 6, "ldb_416x240": 6, "ra_1080p": 32, "ld_1080p_s3001": 9, "ld_1080p_s3002": 9,
           "ra_4k": 17, "intra_8k": 1}


def _md5(path):
    h = hashlib.md5()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 22), b""):
            h.update(blk)
    return h.hexdigest()


@pytest.mark.parametrize("stream", list(STREAMS))
def test_decoder_with_gpu_filters_passes_hash_sei(stream, tmp_path):
    bit = os.path.join(ROOT, "tests", "golden", "streams", stream + ".bin")
    out = str(tmp_path / "gpu.yuv")
    r = subprocess.run([SHIM_DEC, "-b", bit, "-o", out, "-d", "10"], capture_output=True, text=True, env=dict(os.environ, ILF_TIMING="1"))
    assert r.returncode == 0, (r.stdout + r.stderr)[-3000:]
    assert "ERROR" not in r.stdout and "mismatch" not in r.stdout.lower(), r.stdout[-3000:]
    assert r.stdout.count("(OK)") == STREAMS[stream], r.stdout[-3000:]
    assert r.stderr.count("impl=b200") == STREAMS[stream]            # every picture went through the CUDA library
    if os.path.exists(STOCK_DEC):
        ref = str(tmp_path / "cpu.yuv")
        subprocess.run([STOCK_DEC, "-b", bit, "-o", ref, "-d", "10"], check=True, capture_output=True)
        assert _md5(out) == _md5(ref)
