"""Band decomposition on the CPU with the oracle as the per-band filter (test infrastructure for the N>1 path).

The oracle filters whole pictures whose CTU grid starts at row 0, so the CPU emulation of a band uses one CTU row of
halo on each side (the CUDA path holds 16 luma rows, tests/test_gpu_bands.py); what is checked here is the host logic:
partition, neighbour exchange pattern, slicing of the side information, and that a band's own rows do not depend on
anything outside band + halo."""
import numpy as np

import synth

K = ("y", "cb", "cr")


def region_of(band, ctus_h):
    """CTU rows [r0, r1) a rank filters: its band plus one halo CTU row on each side."""
    first, n = band
    return max(0, first - 1), min(ctus_h, first + n + 1)


def filter_region(oracle, planes, r0, r1, bd, ctu_log2, ctus_w, ctus_h, height, db, sao_ctus, alf):
    """deblock -> SAO -> ALF of CTU rows [r0, r1) treated as a picture; `planes` holds exactly those rows."""
    ctu = 1 << ctu_log2
    y0, y1 = r0 * ctu, min(height, r1 * ctu)
    prm, info, mv16 = db
    info = info[y0 // 4:y1 // 4].copy()
    if r0 > 0:
        info[0] &= ~np.uint32(0x30)   # the region's top border is a picture border for the oracle: no horizontal edge there
    mv = None if mv16 is None else np.ascontiguousarray(mv16[y0 // 4:y1 // 4])
    out = oracle.deblock(planes, bd, bd, ctu_log2, prm, info, None, mv, None, None)
    sao = sao_ctus.view(synth.SAO_DT).reshape(ctus_h, ctus_w)[r0:r1].copy()
    if r0 > 0:
        sao["avail"][0] &= ~np.uint8(0x04 | 0x10 | 0x20)
    if r1 < ctus_h:
        sao["avail"][-1] &= ~np.uint8(0x08 | 0x40 | 0x80)
    out = oracle.sao(out, bd, bd, ctu_log2, sao.view(np.uint8).reshape(-1, 32))
    alf_bytes, en = alf
    en = en.reshape(3, ctus_h, ctus_w)[:, r0:r1].reshape(3, -1)
    return oracle.alf(out, bd, bd, ctu_log2, alf_bytes, np.ascontiguousarray(en))


def rows(pic, y0, y1):
    return {"y": pic["y"][y0:y1], "cb": pic["cb"][y0 // 2:y1 // 2], "cr": pic["cr"][y0 // 2:y1 // 2]}
