// ilf_sao.cu -- sample adaptive offset (sm_100a).
//
// Replaces SampleAdaptiveOffset::SAOProcess after parameter resolution (SampleAdaptiveOffset.cpp:585-601,
// offsetCTU :510-562, offsetBlock :292-508).  The source is the whole deblocked picture (the reference copies it to
// m_tempBuf, :587), so every sample is independent.
//
// Work split: one thread owns a strip of 8 samples x 8 rows (one int16x8 vector per row).  It issues all of its
// global loads first (8 rows + the row above and below for the vertical classes: 160 bytes in flight per thread),
// then filters two samples per instruction (ilf_packed.cuh) and stores 8 vectors.  A strip lies inside one CTU, so
// the CTU's parameters are fetched once per thread; the lanes of a warp that share a row group span exactly one CTU
// width at the 128x128 CTU size (16 lanes luma, 8 lanes chroma), so a warp does not diverge on the SAO type.
// Horizontal neighbours come from the adjacent lane by shuffle; the first/last lane of a row group fetches one
// sample per row from the neighbouring CTU.  CTUs with SAO off are copied through; planes whose SAO is off for the
// whole picture are not touched at all (BatchCtl skip bit).
//
// The row/column special cases of offsetBlock (:308-487) are the statement "an edge-offset sample is modified iff both
// neighbours along the class direction are inside the CTU block or inside a neighbouring CTU whose availability flag
// (deriveLoopFilterBoundaryAvailibility, :685-760) is set"; it is evaluated per row for the first column, the last
// column and the columns in between, and only for CTUs that have an unavailable neighbour.
#include "ilf_common.cuh"
#include "ilf_packed.cuh"

namespace ilf {
namespace {

constexpr int R = 8;          // rows per strip
constexpr int NTHREADS = 128;

struct Row { uint32_t v[4]; };  // 8 samples

__device__ __forceinline__ Row ld_row(const int16_t* p) {
  const uint4 r = ldg_u4(p);
  Row o; o.v[0] = r.x; o.v[1] = r.y; o.v[2] = r.z; o.v[3] = r.w;
  return o;
}
__device__ __forceinline__ void st_row(int16_t* p, const Row& r) { *reinterpret_cast<uint4*>(p) = make_uint4(r.v[0], r.v[1], r.v[2], r.v[3]); }

// The row shifted by one sample: s[0] = [x-1, x0], s[1..3] = [x1,x2] [x3,x4] [x5,x6], s[4] = [x7, x+8].
// left neighbours of the 8 samples = s[0..3], right neighbours = s[1..4].
struct Shifted { uint32_t s[5]; };
__device__ __forceinline__ Shifted shift_row(const Row& r, uint32_t left, uint32_t right) {
  Shifted o;
  o.s[0] = (r.v[0] << 16) | (left & 0xFFFFu);
  o.s[1] = pk::funnel16(r.v[0], r.v[1]);
  o.s[2] = pk::funnel16(r.v[1], r.v[2]);
  o.s[3] = pk::funnel16(r.v[2], r.v[3]);
  o.s[4] = (r.v[3] >> 16) | (right << 16);
  return o;
}

// Availability of the region a neighbour falls in.  dxr/dyr in {-1,0,1}: left/inside/right, above/inside/below.
__device__ __forceinline__ bool region_ok(unsigned m9, int dxr, int dyr) { return (m9 >> ((dyr + 1) * 3 + dxr + 1)) & 1u; }

struct EoCtx {
  uint32_t lut_lo, lut_hi, maxv;
  unsigned m9;
  int ctu_y0, last_row, rj, gy0, nrows;
  bool at_left, need_mask;
};

// Edge offset of the strip's rows for one class: first neighbour a = (x - SX, y - SY), second b = (x + SX, y + SY).
// sft[0] = shifted row above the strip, sft[1 + r] = shifted row r, sft[nrows + 1] = shifted row below.
template <int SX, int SY>
__device__ __forceinline__ void eo_rows(const EoCtx& e, const Row (&c)[R], const Row& above, const Row& below, const Shifted (&sft)[R + 2],
                                        int16_t* __restrict__ out, int pitch) {
#pragma unroll
  for (int r = 0; r < R; r++) {
    if (r >= e.nrows) break;
    uint32_t a[4], b[4];
    if (SX == 0) {
      const Row& ra = r == 0 ? above : c[r > 0 ? r - 1 : 0];
      const Row& rb = (r + 1 < e.nrows) ? c[r + 1 < R ? r + 1 : R - 1] : below;
#pragma unroll
      for (int k = 0; k < 4; k++) { a[k] = ra.v[k]; b[k] = rb.v[k]; }
    } else {
      constexpr int ia_off = SY ? 0 : 1, ib_off = SY ? 2 : 1;
#pragma unroll
      for (int k = 0; k < 4; k++) {
        a[k] = sft[r + ia_off].s[SX > 0 ? k : k + 1];
        b[k] = sft[r + ib_off].s[SX > 0 ? k + 1 : k];
      }
    }
    Row o;
#pragma unroll
    for (int k = 0; k < 4; k++) o.v[k] = pk::sao_apply2(c[r].v[k], pk::sao_eo_index2(c[r].v[k], a[k], b[k]), e.lut_lo, e.lut_hi, e.maxv);
    if (e.need_mask) {
      const int gy = e.gy0 + r;
      const bool at_top = gy == e.ctu_y0, at_bottom = gy == e.last_row;
      const int dy1 = (at_top && SY) ? -1 : 0, dy2 = (at_bottom && SY) ? 1 : 0;
      const bool ok_mid = region_ok(e.m9, 0, dy1) && region_ok(e.m9, 0, dy2);
      const bool ok_first = region_ok(e.m9, SX > 0 ? -1 : 0, dy1) && region_ok(e.m9, SX < 0 ? -1 : 0, dy2);
      const bool ok_last = region_ok(e.m9, SX < 0 ? 1 : 0, dy1) && region_ok(e.m9, SX > 0 ? 1 : 0, dy2);
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const bool k0 = (k == 0 && e.at_left) ? ok_first : ok_mid;
        const bool l0 = e.rj == 2 * k ? ok_last : k0;
        const bool l1 = e.rj == 2 * k + 1 ? ok_last : ok_mid;
        const uint32_t m = (l0 ? 0xFFFFu : 0u) | (l1 ? 0xFFFF0000u : 0u);
        o.v[k] = (o.v[k] & m) | (c[r].v[k] & ~m);
      }
    }
    st_row(out + (size_t)r * pitch, o);
  }
}

template <int LPR>
__device__ __forceinline__ void sao_strip(const Geom& g, const SlotDev& sd, unsigned ctl, int plane, int bx, int by) {
  constexpr int RG = 32 / LPR;  // row groups per warp
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int seg_lane = lane % LPR, rgrp = lane / LPR;
  const int sh = plane ? 1 : 0;
  const int pw = g.width >> sh, ph_local = g.rows >> sh, ph_global = g.height >> sh;
  const int pitch = plane ? g.pitch_c : g.pitch_y;
  const int x0 = (bx * LPR + seg_lane) * 8;
  const int y0 = ((by * (NTHREADS / 32) + warp) * RG + rgrp) * R;  // local row of the strip's first row
  const bool in = x0 < pw && y0 < ph_local;
  const unsigned full = 0xffffffffu;

  const int16_t* __restrict__ src = sd.buf[ctl_src(ctl, plane)][plane];
  int16_t* __restrict__ dst = sd.buf[ctl_dst(ctl, plane)][plane];
  const int nrows = in ? min(R, ph_local - y0) : 0;
  const int gy0 = y0 + (g.row0 >> sh);  // picture row
  const int ctu_log2 = g.ctu_log2 - sh, ctu_sz = 1 << ctu_log2;

  int type = ILF_SAO_OFF;
  uint4 pa = make_uint4(0, 0, 0, 0), pb = pa;
  int cx = 0, cy = 0;
  if (in) {
    cx = x0 >> ctu_log2; cy = gy0 >> ctu_log2;
    const uint4* __restrict__ p = reinterpret_cast<const uint4*>(sd.sao + (size_t)cy * g.ctus_w + cx);
    pa = __ldg(p); pb = __ldg(p + 1);
    // ilf_sao_ctu: offset[3][4] int16 (24 bytes), type[3] int8 at byte 24, band_pos[3] at 27, avail at 30
    type = (int)(int8_t)((plane == 0 ? pb.z : plane == 1 ? pb.z >> 8 : pb.z >> 16) & 0xFF);
  }
  const bool vert = type == ILF_SAO_EO_90 || type == ILF_SAO_EO_135 || type == ILF_SAO_EO_45;
  const bool horz = type == ILF_SAO_EO_0 || type == ILF_SAO_EO_135 || type == ILF_SAO_EO_45;

  // ---- all loads first ----
  const int16_t* base = src + (size_t)y0 * pitch + x0;
  Row c[R];
#pragma unroll
  for (int r = 0; r < R; r++) if (r < nrows) c[r] = ld_row(base + (ptrdiff_t)r * pitch);
  Row above, below;
  // a lane's rows above/below also feed its neighbours' diagonal taps, so load them when any lane of the warp needs them
  // (the lanes of a row group share one CTU and one type at the 128x128 CTU size; smaller CTUs mix types in a warp)
  const bool any_vert = __any_sync(full, vert);
  if (any_vert && in) {
    above = ld_row(base - (y0 > 0 ? pitch : 0));
    below = ld_row(base + (ptrdiff_t)(y0 + nrows < ph_local ? nrows : nrows - 1) * pitch);
  }
  // halo samples of the first / last lane of a row group (neighbouring CTU column), rows -1 .. R
  uint32_t hl[R + 2], hr[R + 2];
  const bool any_horz = __any_sync(full, horz);
  if (any_horz) {
    const bool edge_l = seg_lane == 0 && horz && x0 > 0, edge_r = seg_lane == LPR - 1 && horz && x0 + 8 < pw;
#pragma unroll
    for (int r = -1; r <= R; r++) {
      hl[r + 1] = 0; hr[r + 1] = 0;
      const bool row_ok = (r >= 0 && r < nrows) || (vert && (r == -1 ? y0 > 0 : (r == nrows && y0 + nrows < ph_local)));  // `horz && vert` lanes only
      if (edge_l && row_ok) hl[r + 1] = (uint16_t)__ldg(base + (ptrdiff_t)r * pitch - 1);
      if (edge_r && row_ok) hr[r + 1] = (uint16_t)__ldg(base + (ptrdiff_t)r * pitch + 8);
    }
  }
  if (__all_sync(full, !in)) return;

  if (type == ILF_SAO_OFF && !any_horz) {  // plain copy
#pragma unroll
    for (int r = 0; r < R; r++) if (r < nrows) st_row(dst + (size_t)(y0 + r) * pitch + x0, c[r]);
    return;
  }

  // ---- neighbour exchange inside the row group (all lanes take part) ----
  // rows: index 0 = above, 1..R = c[0..R-1], R+1 = below (the row after the strip's last row)
  Shifted sft[R + 2];
  if (any_horz) {
#pragma unroll
    for (int i = 0; i < R + 2; i++) {
      Row rr;
      if (i == 0) rr = above;
      else if (i == R + 1) rr = below;
      else rr = c[i - 1];
      if (i >= 1 && i <= R && i - 1 == nrows) rr = below;  // short strip at the picture bottom: "below" follows the last valid row
      const uint32_t lft = __shfl_up_sync(full, rr.v[3] >> 16, 1, LPR);
      const uint32_t rgt = __shfl_down_sync(full, rr.v[0] & 0xFFFFu, 1, LPR);
      sft[i] = shift_row(rr, seg_lane == 0 ? hl[i] : lft, seg_lane == LPR - 1 ? hr[i] : rgt);
    }
  }
  if (!in) return;
  if (type == ILF_SAO_OFF) {
#pragma unroll
    for (int r = 0; r < R; r++) if (r < nrows) st_row(dst + (size_t)(y0 + r) * pitch + x0, c[r]);
    return;
  }

  // ---- parameters of this component ----
  const uint32_t w0 = plane == 0 ? pa.x : plane == 1 ? pa.z : pb.x;  // offset[plane][0..1]
  const uint32_t w1 = plane == 0 ? pa.y : plane == 1 ? pa.w : pb.y;  // offset[plane][2..3]
  const uint32_t o0 = w0 & 0xFF, o1 = (w0 >> 16) & 0xFF, o2 = w1 & 0xFF, o3 = (w1 >> 16) & 0xFF;  // int8 range checked on the host
  const int bd = plane ? g.bd_chroma : g.bd_luma;
  const uint32_t maxv = pk::splat((1 << bd) - 1);
  const unsigned avail = (pb.w >> 16) & 0xFF;

  if (type == ILF_SAO_BO) {
    const uint32_t lut_lo = o0 | (o1 << 8) | (o2 << 16) | (o3 << 24), lut_hi = 0;
    const int band = (plane == 0 ? pb.z >> 24 : plane == 1 ? pb.w : pb.w >> 8) & 0xFF;
    const uint32_t nband = pk::splat((32 - band) & 31);
#pragma unroll
    for (int r = 0; r < R; r++) {
      if (r >= nrows) break;
      Row o;
#pragma unroll
      for (int k = 0; k < 4; k++) o.v[k] = pk::sao_apply2(c[r].v[k], pk::sao_bo_index2(c[r].v[k], bd - 5, nband), lut_lo, lut_hi, maxv);
      st_row(dst + (size_t)(y0 + r) * pitch + x0, o);
    }
    return;
  }

  EoCtx e;
  e.lut_lo = o0 | (o1 << 8) | (o2 << 24); e.lut_hi = o3;  // index 2 (flat) adds 0
  e.maxv = maxv;
  // 3x3 availability mask of the CTU: bit (dyr+1)*3 + (dxr+1)
  e.m9 = ((avail & ILF_AVAIL_AL) ? 1u : 0u) | ((avail & ILF_AVAIL_A) ? 2u : 0u) | ((avail & ILF_AVAIL_AR) ? 4u : 0u) |
         ((avail & ILF_AVAIL_L) ? 8u : 0u) | 16u | ((avail & ILF_AVAIL_R) ? 32u : 0u) |
         ((avail & ILF_AVAIL_BL) ? 64u : 0u) | ((avail & ILF_AVAIL_B) ? 128u : 0u) | ((avail & ILF_AVAIL_BR) ? 256u : 0u);
  const int ctu_x0 = cx << ctu_log2;
  e.ctu_y0 = cy << ctu_log2;
  e.at_left = x0 == ctu_x0;
  e.rj = min(ctu_x0 + ctu_sz, pw) - x0 - 1;  // index of the block's last column inside this vector (>= 8: not here)
  e.last_row = min(e.ctu_y0 + ctu_sz, ph_global) - 1;
  e.need_mask = avail != 0xFFu;
  e.gy0 = gy0; e.nrows = nrows;
  int16_t* out = dst + (size_t)y0 * pitch + x0;
  if (type == ILF_SAO_EO_0) eo_rows<1, 0>(e, c, above, below, sft, out, pitch);
  else if (type == ILF_SAO_EO_90) eo_rows<0, 1>(e, c, above, below, sft, out, pitch);
  else if (type == ILF_SAO_EO_135) eo_rows<1, 1>(e, c, above, below, sft, out, pitch);
  else eo_rows<-1, 1>(e, c, above, below, sft, out, pitch);
}

__global__ void __launch_bounds__(NTHREADS) sao_kernel(Geom g, const SlotDev* __restrict__ slots, int first_slot, BatchCtl bc,
                                                      int gy_y, int gy_c, int gx_c) {
  const unsigned ctl = bc.v[blockIdx.z];
  const SlotDev& sd = slots[first_slot + bc.slot[blockIdx.z]];
  // blockIdx.y enumerates row groups of Y, then Cb, then Cr.
  int by = blockIdx.y;
  if (by < gy_y) {
    if (ctl_skip(ctl, 0)) return;
    sao_strip<16>(g, sd, ctl, 0, blockIdx.x, by);
  } else {
    by -= gy_y;
    const int plane = by < gy_c ? 1 : 2;
    if (plane == 2) by -= gy_c;
    if (ctl_skip(ctl, plane) || (int)blockIdx.x >= gx_c) return;
    sao_strip<8>(g, sd, ctl, plane, blockIdx.x, by);
  }
}

}  // namespace

void launch_sao(const Geom& g, const SlotDev* slots, int first_slot, int num_slots, const BatchCtl& ctl, cudaStream_t st) {
  // luma CTA: 128 samples x 64 rows (4 warps x 2 row groups x 8 rows); chroma CTA: 64 samples x 128 rows
  const int gx_y = (g.width + 127) / 128, gy_y = (g.rows + 63) / 64;
  const int gx_c = (g.width / 2 + 63) / 64, gy_c = (g.rows / 2 + 127) / 128;
  dim3 grid(gx_y > gx_c ? gx_y : gx_c, gy_y + 2 * gy_c, num_slots);
  sao_kernel<<<grid, NTHREADS, 0, st>>>(g, slots, first_slot, ctl, gy_y, gy_c, gx_c);
}

}  // namespace ilf
