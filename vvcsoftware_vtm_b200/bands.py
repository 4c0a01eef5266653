"""Host-side sharding of the in-loop filter path across GPUs (DESIGN.md "multi-GPU").

Two natural shards, neither needs a collective on the data path:
  * independent pictures / streams: dealt round-robin to the ranks (`deal_streams`);
  * one picture in CTU-row bands: contiguous CTU rows per rank (`band_partition`), halo rows pulled once per picture
    from the neighbouring ranks' input planes over NVLink P2P (`connect_bands` + InLoopFilter.band_exchange).
"""
import numpy as np

from .ilf import BAND_ABOVE, BAND_BELOW, InLoopFilter


def band_partition(ctu_rows, n):
    """[(first_ctu_row, num_ctu_rows)] for n bands: contiguous, the first (ctu_rows % n) bands one row taller
    (34 CTU rows -> 17/17, 9/9/8/8, 5/5/4/4/4/4/4/4)."""
    if n < 1 or n > ctu_rows:
        raise ValueError(f"cannot split {ctu_rows} CTU rows into {n} bands")
    base, extra = divmod(ctu_rows, n)
    out, first = [], 0
    for i in range(n):
        rows = base + (1 if i < extra else 0)
        out.append((first, rows))
        first += rows
    return out


def deal_streams(num_streams, world):
    """Stream indices per rank, round-robin (64 streams on 8 GPUs -> rank r takes r, r + 8, ...)."""
    return [list(range(r, num_streams, world)) for r in range(world)]


def band_rows(height, ctu_log2, band, halo=16):
    """(own_first, own_rows, held_first, held_rows) in luma rows of a band (first_ctu_row, num_ctu_rows) -- what
    ilf_create_band computes (csrc/ilf_api.cu create_impl)."""
    first, n = band
    own0 = first << ctu_log2
    own1 = min(height, (first + n) << ctu_log2)
    held0 = max(0, own0 - halo)
    held1 = min(height, own1 + halo)
    return own0, own1 - own0, held0, held1 - held0


def slice_side_info(si, held_first, held_rows):
    """Side information of a band context: unit grids cover the held rows, everything else the full picture."""
    u0, u1 = held_first // 4, (held_first + held_rows) // 4
    d = dict(si)
    for k in ("db_info", "db_info_c", "db_mv16", "db_mv32"):
        if d.get(k) is not None:
            d[k] = np.ascontiguousarray(d[k][u0:u1])
    return d


def connect_bands(ctx, slot, rank, world, handles):
    """Connect band context `ctx` of rank `rank` to its neighbours; handles[r] = band_export() of rank r's context."""
    if rank > 0:
        ctx.band_connect(slot, BAND_ABOVE, handles[rank - 1])
    if rank < world - 1:
        ctx.band_connect(slot, BAND_BELOW, handles[rank + 1])


def make_band_context(width, height, bd_luma, bd_chroma, ctu_log2, rank, world, device, num_slots=1):
    ctu = 1 << ctu_log2
    band = band_partition((height + ctu - 1) // ctu, world)[rank]
    return InLoopFilter(width, height, bd_luma, bd_chroma, ctu_log2, device=device, num_slots=num_slots, band=band)
