"""Host side of the multi-GPU path on the CPU: sharding helpers, and a world_size-2/3 gloo run in which every rank filters
its CTU-row band (oracle as the per-band filter) after receiving its halo from the neighbouring ranks."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import band_util as BU
import synth
from vvcsoftware_vtm_b200 import bands


def test_band_partition_matches_survey():
    assert bands.band_partition(34, 2) == [(0, 17), (17, 17)]
    assert [n for _, n in bands.band_partition(34, 4)] == [9, 9, 8, 8]
    assert [n for _, n in bands.band_partition(34, 8)] == [5, 5, 4, 4, 4, 4, 4, 4]
    for rows in (1, 2, 9, 17, 34):
        for n in range(1, min(rows, 8) + 1):
            p = bands.band_partition(rows, n)
            assert p[0][0] == 0 and sum(k for _, k in p) == rows
            assert all(p[i][0] + p[i][1] == p[i + 1][0] for i in range(n - 1))
    with pytest.raises(ValueError):
        bands.band_partition(2, 3)


def test_deal_streams_round_robin():
    d = bands.deal_streams(64, 8)
    assert d[0][:3] == [0, 8, 16] and all(len(x) == 8 for x in d)
    assert sorted(sum(bands.deal_streams(10, 4), [])) == list(range(10))


def test_band_rows_and_side_info_slices():
    # 1080p, CTU 128: 9 CTU rows, 3 bands of 3; last CTU row is 56 rows high
    p = bands.band_partition(9, 3)
    assert bands.band_rows(1080, 7, p[0]) == (0, 384, 0, 400)
    assert bands.band_rows(1080, 7, p[1]) == (384, 384, 368, 416)
    assert bands.band_rows(1080, 7, p[2]) == (768, 312, 752, 328)
    si = {"db_info": np.arange(270 * 4, dtype=np.uint32).reshape(270, 4), "db_mv16": None, "sao_ctus": np.zeros((135, 32), np.uint8)}
    s = bands.slice_side_info(si, 368, 416)
    assert s["db_info"].shape == (104, 4) and s["db_info"][0, 0] == 92 * 4 and s["sao_ctus"] is si["sao_ctus"]


W, H, BD, CTU_LOG2 = 200, 136, 10, 5   # 7 x 5 CTUs of 32 x 32


def _case(seed):
    rng = np.random.default_rng(seed)
    ctu = 1 << CTU_LOG2
    cw, ch = (W + ctu - 1) // ctu, (H + ctu - 1) // ctu
    pic = synth.picture(rng, W, H, BD)
    db = synth.deblock_info(rng, W, H, inter=True)
    sao = synth.sao_params(rng, cw, ch, BD, p_off=0.1)
    alf = synth.alf_params(rng, cw, ch, is7=True, p_on=0.9)
    return pic, db, sao, alf, cw, ch


def _worker(rank, world, port, seed, q):
    sys.path[:0] = [os.path.dirname(os.path.abspath(__file__)), os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle")]
    import ilf_oracle as O
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        pic, db, sao, alf, cw, ch = _case(seed)   # every rank knows the side information; only its OWN rows of the picture
        ctu = 1 << CTU_LOG2
        part = bands.band_partition(ch, world)
        own0 = part[rank][0] * ctu
        own1 = min(H, (part[rank][0] + part[rank][1]) * ctu)
        mine = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in BU.rows(pic, own0, own1).items()}
        # halo exchange: every rank sends its first CTU row up and its last CTU row down, and receives the matching rows
        def edge_rows(r, which):   # picture rows of rank r's first / last CTU row
            f, n = part[r]
            lo = f * ctu if which == "first" else (f + n - 1) * ctu
            return lo, min(H, lo + ctu)
        halo, reqs, keep = {}, [], []
        for nb, send_which, recv_which, key in ((rank - 1, "first", "last", "above"), (rank + 1, "last", "first", "below")):
            if nb < 0 or nb >= world:
                continue
            s0, s1 = edge_rows(rank, send_which)
            g0, g1 = edge_rows(nb, recv_which)
            halo[key] = {}
            for k in BU.K:
                sh = 0 if k == "y" else 1
                src = mine[k][(s0 - own0) >> sh:(s1 - own0) >> sh].contiguous()
                keep.append(src)
                reqs.append(dist.isend(src, nb))
                buf = torch.empty(((g1 - g0) >> sh, W >> sh), dtype=torch.int16)
                reqs.append(dist.irecv(buf, nb))
                halo[key][k] = buf
        for r in reqs:
            r.wait()
        r0, r1 = BU.region_of(part[rank], ch)
        region = {k: np.concatenate([a.numpy() for a in ([halo["above"][k]] if "above" in halo else []) + [mine[k]] + ([halo["below"][k]] if "below" in halo else [])]) for k in BU.K}
        out = BU.filter_region(O, region, r0, r1, BD, CTU_LOG2, cw, ch, H, db, sao, alf)
        off = own0 - r0 * ctu
        own = BU.rows(out, off, off + (own1 - own0))
        gathered = [None] * world
        dist.all_gather_object(gathered, {k: v.copy() for k, v in own.items()})
        if rank == 0:
            whole = BU.filter_region(O, pic, 0, ch, BD, CTU_LOG2, cw, ch, H, db, sao, alf)
            got = {k: np.concatenate([g[k] for g in gathered]) for k in BU.K}
            q.put({k: int((got[k] != whole[k]).sum()) for k in BU.K})
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,seed", [(2, 11), (3, 12)])
def test_banded_filtering_over_gloo_equals_whole_picture(world, seed, oracle):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + world * 7 + seed
    procs = [ctx.Process(target=_worker, args=(r, world, port, seed, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0, "a rank failed"
    diff = q.get(timeout=10)
    assert not any(diff.values()), f"banded result differs from the whole-picture result: {diff}"
