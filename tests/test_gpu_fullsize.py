"""Full-size parity (BASELINE configs 2-5 geometries): the CUDA chain on 1080p / 4K / 8K pictures with the REAL side
information of the committed workload streams (bench_data/*.npz, packed from the streams by the product packer) must equal
the oracle sample for sample -- with the stream's own SAO/ALF decisions and with every CTU of every component switched on --
and a device-resident batch must equal its pictures filtered one by one (checksum of checksums over the batch)."""
import os
import sys
import zlib

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (side-information loader and the synthetic planes of the benchmark)
from vvcsoftware_vtm_b200 import bands  # noqa: E402

pytestmark = pytest.mark.gpu
K = ("y", "cb", "cr")
BD, CTU = 10, 7


def _set(f, slot, si):
    f.set_deblock_info(slot, si["db_params"].tobytes(), si["db_info"], si.get("db_info_c"), si.get("db_mv16"), si.get("db_mv32"), si.get("ctu_slice"))
    f.set_sao_params(slot, si["sao_ctus"])
    f.set_alf_params(slot, si["alf_params"].tobytes(), si["alf_ctu_enable"])


def _oracle_chain(O, pic, si):
    out = O.deblock(pic, BD, BD, CTU, si["db_params"].tobytes(), si["db_info"], si.get("db_info_c"), si.get("db_mv16"), si.get("db_mv32"), si.get("ctu_slice"))
    out = O.sao(out, BD, BD, CTU, si["sao_ctus"])
    return O.alf(out, BD, BD, CTU, si["alf_params"].tobytes(), si["alf_ctu_enable"])


def _crc(pic):
    return zlib.crc32(b"".join(np.ascontiguousarray(pic[k]).tobytes() for k in K))


@pytest.mark.parametrize("workload,w,h,picks", [("ra_1080p", 1920, 1080, (0, 5, 9)), ("ld_1080p_s3002", 1920, 1080, (1, 4)), ("ra_4k", 3840, 2160, (0, 4, 11)),
                                                ("intra_8k", 7680, 4320, (0,))])
@pytest.mark.parametrize("all_on", [False, True], ids=["stream", "all_on"])
def test_chain_equals_oracle_at_full_size(workload, w, h, picks, all_on, ilf_lib, oracle):
    side = bench.load_sideinfo(workload)
    if all_on:
        side = bench.all_on_sideinfo(side)
    planes = bench.synth_planes(w, h, len(picks), seed=w + len(side))
    with ilf_lib.InLoopFilter(w, h, BD, BD, CTU) as f:
        for j, p in enumerate(picks):
            si = side[p % len(side)]
            pic = dict(zip(K, planes[j]))
            f.upload(0, *planes[j])
            _set(f, 0, si)
            f.run(0, 1, 7)
            got = f.download(0)
            want = _oracle_chain(oracle, pic, si)
            bad = {k: int((got[k] != want[k]).sum()) for k in K}
            assert not any(bad.values()), f"{workload} picture {p}: mismatching samples {bad}"


def test_batch_of_4k_pictures_equals_one_by_one(ilf_lib):
    """17 resident 4K pictures (one random-access GOP) through one launch per stage == the same pictures one at a time."""
    side = bench.load_sideinfo("ra_4k")
    n = len(side)
    planes = bench.synth_planes(3840, 2160, 3, seed=77)
    singles = []
    with ilf_lib.InLoopFilter(3840, 2160, BD, BD, CTU) as f:
        for j in range(n):
            f.upload(0, *planes[j % 3])
            _set(f, 0, side[j])
            f.run(0, 1, 7)
            singles.append(_crc(f.download(0)))
    with ilf_lib.InLoopFilter(3840, 2160, BD, BD, CTU, num_slots=n) as f:
        for j in range(n):
            f.upload(j, *planes[j % 3])
            _set(f, j, side[j])
        f.run(0, n, 7)
        batch = [_crc(f.download(j)) for j in range(n)]
    assert batch == singles
    assert zlib.crc32(np.array(batch, np.uint32).tobytes()) == zlib.crc32(np.array(singles, np.uint32).tobytes())


@pytest.mark.parametrize("n", [2, 4, 8])
def test_8k_bands_equal_whole_picture(n, ilf_lib):
    """BASELINE config 4: the 8K intra picture in n CTU-row bands (17/17, 9/9/8/8, 5/5/4/...), halos exchanged on the
    input, every band filtered on its own: the bands' rows, concatenated, are the whole-picture result."""
    w, h = 7680, 4320
    si = bench.all_on_sideinfo(bench.load_sideinfo("intra_8k"))[0]
    pic = dict(zip(K, bench.synth_planes(w, h, 1, seed=8000)[0]))
    with ilf_lib.InLoopFilter(w, h, BD, BD, CTU) as f:
        f.upload(0, *(pic[k] for k in K))
        _set(f, 0, si)
        f.run(0, 1, 7)
        whole = f.download(0)
    part = bands.band_partition((h + 127) // 128, n)
    ctxs = [ilf_lib.InLoopFilter(w, h, BD, BD, CTU, band=b) for b in part]
    try:
        for f in ctxs:
            y0, y1 = f.own_row0, f.own_row0 + f.own_rows
            f.upload_band(0, pic["y"][y0:y1], pic["cb"][y0 // 2:y1 // 2], pic["cr"][y0 // 2:y1 // 2])
        for f in ctxs:
            f.sync()
        handles = [f.band_export(0) for f in ctxs]
        for r, f in enumerate(ctxs):
            bands.connect_bands(f, 0, r, n, handles)
            f.band_exchange(0)
            _set(f, 0, bands.slice_side_info(si, f.row0, f.rows))
            f.run(0, 1, 7)
        outs = [f.download_band(0) for f in ctxs]
    finally:
        for f in ctxs:
            f.close()
    for k in K:
        assert np.array_equal(np.concatenate([o[k] for o in outs]), whole[k]), f"plane {k}: banded != whole picture"
