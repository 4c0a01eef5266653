// ilf_deblock.cu -- deblocking filter, both edge directions fused in one pass (sm_100a).
//
// Replaces LoopFilter::loopFilterPic (LoopFilter.cpp:149-230): "all vertical edges of the picture, then all
// horizontal edges".  Filtered edges lie on the 8x8 luma grid (:313-324), each reads 4 and writes 3 samples
// per side (:856-916), and the on/off + strong/weak decisions of a 4-line segment use lines 0 and 3 of that
// segment only (:640-671).  Therefore the dependency closure of every 8x8 block SHIFTED by (-4,-4) is the
// block itself: vertical-edge filtering of its 8 rows needs only its 8 columns, and the horizontal edge in
// its middle needs only those 8 vertically filtered rows.  A CTA therefore stages one shifted tile
// (64x32 luma + two 32x16 chroma tiles) in shared memory, runs the vertical pass and then the horizontal
// pass on it, and writes it out: every sample is read from HBM once and written once, with no halo.
// Chroma edges lie on the 8x8 chroma grid and reach 2/1 samples, so any shift in [2,6] closes them; the
// chroma tile is shifted by (-4,-2) to keep 8-byte alignment of its rows.
//
// Per-edge derivation on the device (xGetBoundaryStrengthSingle :419-541, QP/tc/beta :626-634, chroma QP
// :811-829) from the packed per-4x4 grid described in include/ilf_b200.h.  The 4 lanes that hold the 4 lines
// of a segment exchange their second-derivative terms with warp shuffles.
#include "ilf_common.cuh"

namespace ilf {
namespace {

constexpr int TW = 64, TH = 32;         // luma tile
constexpr int LP = TW + 8;              // luma smem pitch (samples); 144 B rows keep 16-B alignment
constexpr int CTW = 32, CTH = 16;       // chroma tile
constexpr int CP = CTW + 8;             // chroma smem pitch
constexpr int MW = 18, MH = 8;          // staged metadata window in units
constexpr int NTHREADS = 256;

__constant__ uint8_t c_tc[66] = {0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  1,  1,  1,  1,
                                 1,  1,  1,  1,  1,  2,  2,  2,  2,  3,  3,  3,  3,  4,  4,  4,  5,  5,  6,  6,  7,  8,
                                 9,  10, 11, 13, 14, 16, 18, 20, 22, 24, 26, 28, 30, 32, 34, 36, 38, 40, 42, 44, 46, 48};
__constant__ uint8_t c_beta[64] = {0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  6,  7,  8,  9,  10, 11,
                                   12, 13, 14, 15, 16, 17, 18, 20, 22, 24, 26, 28, 30, 32, 34, 36, 38, 40, 42, 44, 46, 48,
                                   50, 52, 54, 56, 58, 60, 62, 64, 66, 68, 70, 72, 74, 76, 78, 80, 82, 84, 86, 88};
__constant__ uint8_t c_chroma_scale[70] = {0,  1,  2,  3,  4,  5,  6,  7,  8,  9,  10, 11, 12, 13, 14, 15, 16, 17,
                                           18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 29, 30, 31, 32, 33, 33,
                                           34, 34, 35, 35, 36, 36, 37, 37, 38, 39, 40, 41, 42, 43, 44, 45, 46, 47,
                                           48, 49, 50, 51, 52, 53, 54, 55, 56, 57, 58, 59, 60, 61, 62, 63};

struct Smem {
  int16_t y[TH * LP];
  int16_t c[2][CTH * CP];
  uint32_t info[MH * MW];
  uint32_t info_c[MH * MW];
  int mv[MH * MW * 4];
};

struct EdgeParams {
  int bs;        // 0 = leave the segment alone
  int tc, beta;
  bool no_p, no_q;
};

// bS of one 4-sample segment.  q, p index the staged metadata window.
template <int MV>
__device__ __forceinline__ int boundary_strength(const Smem& s, const uint32_t* info, int q, int p, uint32_t tu_bit,
                                                 int thr) {
  const uint32_t iq = info[q], ip = info[p];
  if ((iq | ip) & ILF_BI_INTRA) return 2;
  if ((iq & tu_bit) && ((iq | ip) & ILF_BI_CBF)) return 1;
  const int rq0 = (iq >> 16) & 0xFF, rq1 = iq >> 24, rp0 = (ip >> 16) & 0xFF, rp1 = ip >> 24;
  const int* mq = &s.mv[q * 4];
  const int* mp = &s.mv[p * 4];
  const bool d00 = abs(mq[0] - mp[0]) >= thr || abs(mq[1] - mp[1]) >= thr;
  if ((iq | ip) & ILF_BI_BSLICE) {
    if ((rp0 == rq0 && rp1 == rq1) || (rp0 == rq1 && rp1 == rq0)) {
      const bool d11 = abs(mq[2] - mp[2]) >= thr || abs(mq[3] - mp[3]) >= thr;
      const bool d10 = abs(mq[2] - mp[0]) >= thr || abs(mq[3] - mp[1]) >= thr;
      const bool d01 = abs(mq[0] - mp[2]) >= thr || abs(mq[1] - mp[3]) >= thr;
      if (rp0 != rp1) return (rp0 == rq0) ? (d00 || d11) : (d10 || d01);
      return (d00 || d11) && (d10 || d01);
    }
    return 1;
  }
  if (rp0 != rq0) return 1;
  return d00;
}

// Luma filter of one line across an edge.  v[0..7] = p3 p2 p1 p0 | q0 q1 q2 q3 (m0..m7 of xPelFilterLuma).
__device__ __forceinline__ void filter_luma_line(int v[8], int tc, bool sw, bool no_p, bool no_q, int thr_cut,
                                                 bool second_p, bool second_q, int max_val) {
  const int m0 = v[0], m1 = v[1], m2 = v[2], m3 = v[3], m4 = v[4], m5 = v[5], m6 = v[6], m7 = v[7];
  if (sw) {
    const int t2 = 2 * tc;
    v[3] = clip3i(m3 - t2, m3 + t2, (m1 + 2 * m2 + 2 * m3 + 2 * m4 + m5 + 4) >> 3);
    v[4] = clip3i(m4 - t2, m4 + t2, (m2 + 2 * m3 + 2 * m4 + 2 * m5 + m6 + 4) >> 3);
    v[2] = clip3i(m2 - t2, m2 + t2, (m1 + m2 + m3 + m4 + 2) >> 2);
    v[5] = clip3i(m5 - t2, m5 + t2, (m3 + m4 + m5 + m6 + 2) >> 2);
    v[1] = clip3i(m1 - t2, m1 + t2, (2 * m0 + 3 * m1 + m2 + m3 + m4 + 4) >> 3);
    v[6] = clip3i(m6 - t2, m6 + t2, (m3 + m4 + m5 + 3 * m6 + 2 * m7 + 4) >> 3);
  } else {
    int delta = (9 * (m4 - m3) - 3 * (m5 - m2) + 8) >> 4;
    if (abs(delta) < thr_cut) {
      delta = clip3i(-tc, tc, delta);
      v[3] = clip3i(0, max_val, m3 + delta);
      v[4] = clip3i(0, max_val, m4 - delta);
      const int tc2 = tc >> 1;
      if (second_p) v[2] = clip3i(0, max_val, m2 + clip3i(-tc2, tc2, (((m1 + m3 + 1) >> 1) - m2 + delta) >> 1));
      if (second_q) v[5] = clip3i(0, max_val, m5 + clip3i(-tc2, tc2, (((m6 + m4 + 1) >> 1) - m5 - delta) >> 1));
    }
  }
  if (no_p) { v[3] = m3; v[2] = m2; v[1] = m1; }
  if (no_q) { v[4] = m4; v[5] = m5; v[6] = m6; }
}

// Decision + filtering of one luma line given the params of its segment.  `l0`/`l3` are the lanes that hold
// lines 0 and 3 of the segment.  All 32 lanes must call this (shuffles).
__device__ __forceinline__ bool luma_line(int v[8], const EdgeParams& ep, int l0, int l3, int max_val) {
  const int dp = abs(v[1] - 2 * v[2] + v[3]);
  const int dq = abs(v[4] - 2 * v[5] + v[6]);
  const int dsum = dp + dq;
  // xUseStrongFiltering on the own line with d = 2 * (dp + dq) of the own line (:670-671, :960-970)
  const bool strong_own = (abs(v[0] - v[3]) + abs(v[7] - v[4]) < (ep.beta >> 3)) && (2 * dsum < (ep.beta >> 2)) &&
                          (abs(v[3] - v[4]) < ((ep.tc * 5 + 1) >> 1));
  const unsigned full = 0xffffffffu;
  const int dp0 = __shfl_sync(full, dp, l0), dp3 = __shfl_sync(full, dp, l3);
  const int dq0 = __shfl_sync(full, dq, l0), dq3 = __shfl_sync(full, dq, l3);
  const unsigned strong_mask = __ballot_sync(full, strong_own);
  if (ep.bs == 0) return false;
  const int d = dp0 + dq0 + dp3 + dq3;
  if (d >= ep.beta) return false;
  const int side_thr = (ep.beta + (ep.beta >> 1)) >> 3;
  const bool sw = ((strong_mask >> l0) & 1u) && ((strong_mask >> l3) & 1u);
  filter_luma_line(v, ep.tc, sw, ep.no_p, ep.no_q, ep.tc * 10, (dp0 + dp3) < side_thr, (dq0 + dq3) < side_thr, max_val);
  return true;
}

template <int MV>
__device__ __forceinline__ EdgeParams luma_edge_params(const Smem& s, const ilf_deblock_params* __restrict__ prm, int q,
                                                       int p, bool vertical, int slice, int bd, bool valid) {
  EdgeParams ep;
  ep.bs = 0; ep.tc = 0; ep.beta = 0; ep.no_p = ep.no_q = false;
  if (!valid) return ep;
  const uint32_t iq = s.info[q], ip = s.info[p];
  if (!(iq & (vertical ? ILF_BI_EDGE_V : ILF_BI_EDGE_H))) return ep;
  const int bs = boundary_strength<MV>(s, s.info, q, p, vertical ? ILF_BI_TU_V : ILF_BI_TU_H, prm->mv_threshold);
  if (!bs) return ep;
  const int qp = ((int)(int8_t)(ip >> 8) + (int)(int8_t)(iq >> 8) + 1) >> 1;
  const int tc_off = prm->slices[slice].tc_offset_div2, beta_off = prm->slices[slice].beta_offset_div2;
  const int scale = 1 << (bd - 8);
  ep.bs = bs;
  ep.tc = c_tc[clip3i(0, 65, qp + 2 * (bs - 1) + 2 * tc_off)] * scale;
  ep.beta = c_beta[clip3i(0, 63, qp + 2 * beta_off)] * scale;
  ep.no_p = (ip & ILF_BI_NOFILT) != 0;
  ep.no_q = (iq & ILF_BI_NOFILT) != 0;
  return ep;
}

// tc of a chroma segment for one component, or -1 when the segment is not filtered (:684-838).
__device__ __forceinline__ int chroma_tc(const uint32_t* info, const ilf_deblock_params* __restrict__ prm, int q, int p,
                                         bool vertical, int slice, int comp, int bd, bool& no_p, bool& no_q) {
  const uint32_t iq = info[q], ip = info[p];
  if (!(iq & (vertical ? ILF_BI_EDGE_V : ILF_BI_EDGE_H))) return -1;
  if (!((iq | ip) & ILF_BI_INTRA)) return -1;  // chroma is filtered for bS == 2 only (:769)
  int qp = (((int)(int8_t)(ip >> 8) + (int)(int8_t)(iq >> 8) + 1) >> 1) + (comp == 0 ? prm->cb_qp_offset : prm->cr_qp_offset);
  if (qp >= 70) qp -= 6;
  else if (qp >= 0) qp = c_chroma_scale[qp];
  no_p = (ip & ILF_BI_NOFILT) != 0;
  no_q = (iq & ILF_BI_NOFILT) != 0;
  return c_tc[clip3i(0, 65, qp + 2 + 2 * prm->slices[slice].tc_offset_div2)] * (1 << (bd - 8));
}

template <int MV>
__global__ void __launch_bounds__(NTHREADS) deblock_kernel(Geom g, const SlotDev* __restrict__ slots, int first_slot, int src_b, int dst_b) {
  __shared__ __align__(16) Smem s;
  const SlotDev& sd = slots[first_slot + blockIdx.z];
  const int tid = threadIdx.x;
  const int tx = blockIdx.x, ty = blockIdx.y;
  const int x0 = tx * TW - 4, y0 = ty * TH - 4;      // luma tile origin, rows are LOCAL (held-region) rows
  const int cx0 = tx * CTW - 4, cy0 = ty * CTH - 2;  // chroma tile origin
  const int ux0 = tx * 16 - 2, uy0 = ty * 8 - 1;     // metadata window origin (units)
  const int16_t* __restrict__ src_y = sd.buf[src_b][0];
  int16_t* __restrict__ dst_y = sd.buf[dst_b][0];
  const int rows = g.rows, crow = g.rows >> 1, cw = g.width >> 1;
  const int units_h_local = rows >> 2;

  // ---- stage: luma tile, 8-byte chunks (the -4 shift keeps 8-byte alignment) ----
#pragma unroll
  for (int i = 0; i < (TH * TW / 4) / NTHREADS; i++) {
    const int c = tid + i * NTHREADS, r = c >> 4, k = c & 15;
    const int x = x0 + k * 4, y = y0 + r;
    if (x >= 0 && x < g.width && y >= 0 && y < rows)
      *reinterpret_cast<uint2*>(&s.y[r * LP + k * 4]) = ldg_u2(src_y + (size_t)y * g.pitch_y + x);
  }
  // ---- chroma tiles: 2 planes x 16 rows x 8 chunks = 256 chunks ----
  {
    const int pl = tid >> 7, c = tid & 127, r = c >> 3, k = c & 7;
    const int x = cx0 + k * 4, y = cy0 + r;
    if (x >= 0 && x < cw && y >= 0 && y < crow)
      *reinterpret_cast<uint2*>(&s.c[pl][r * CP + k * 4]) =
          ldg_u2(sd.buf[src_b][1 + pl] + (size_t)y * g.pitch_c + x);
  }
  // ---- metadata window ----
  const bool has_ctree = sd.info_c != nullptr;
  if (tid < MH * MW) {
    const int mx = tid % MW, my = tid / MW, ux = ux0 + mx, uy = uy0 + my;
    const bool in = ux >= 0 && ux < g.units_w && uy >= 0 && uy < units_h_local;
    const size_t u = (size_t)uy * g.units_w + ux;
    s.info[tid] = in ? __ldg(sd.info + u) : 0u;
    if (has_ctree) s.info_c[tid] = in ? __ldg(sd.info_c + u) : 0u;
    if (MV == 1) {
      uint2 m = in ? ldg_u2(sd.mv16 + u * 4) : make_uint2(0u, 0u);
      s.mv[tid * 4 + 0] = (int)(int16_t)(m.x & 0xFFFF); s.mv[tid * 4 + 1] = (int)(int16_t)(m.x >> 16);
      s.mv[tid * 4 + 2] = (int)(int16_t)(m.y & 0xFFFF); s.mv[tid * 4 + 3] = (int)(int16_t)(m.y >> 16);
    } else if (MV == 2) {
      uint4 m = in ? ldg_u4(sd.mv32 + u * 4) : make_uint4(0u, 0u, 0u, 0u);
      s.mv[tid * 4 + 0] = (int)m.x; s.mv[tid * 4 + 1] = (int)m.y; s.mv[tid * 4 + 2] = (int)m.z; s.mv[tid * 4 + 3] = (int)m.w;
    } else {  // no MV array given: every vector is zero, reference-picture ids still count
      s.mv[tid * 4 + 0] = 0; s.mv[tid * 4 + 1] = 0; s.mv[tid * 4 + 2] = 0; s.mv[tid * 4 + 3] = 0;
    }
  }
  __syncthreads();

  const ilf_deblock_params* __restrict__ prm = sd.db_params;
  const uint32_t* cinfo = has_ctree ? s.info_c : s.info;
  const int max_y = (1 << g.bd_luma) - 1, max_c = (1 << g.bd_chroma) - 1;
  const int lane = tid & 31, warp = tid >> 5;
  const int ctu_row0 = g.row0;  // CTU rows are global

  // ---- vertical edges, luma: warp = the 4 lines of one unit row x 8 edges; lane = line * 8 + edge ----
  {
    const int e = lane & 7, line = lane >> 3;
    const int r = warp * 4 + line;               // tile row
    const int mx = 2 + 2 * e, my = warp;         // Q unit in the metadata window (x = 64tx + 8e -> unit 16tx + 2e)
    const int xg = x0 + 4 + 8 * e, yg = y0 + r;  // first Q sample
    const bool valid = xg > 0 && xg < g.width && yg >= 0 && yg < rows;
    const int slice = (valid && sd.ctu_slice) ? sd.ctu_slice[(size_t)((yg + ctu_row0) >> g.ctu_log2) * g.ctus_w + (xg >> g.ctu_log2)] : 0;
    const EdgeParams ep = luma_edge_params<MV>(s, prm, my * MW + mx, my * MW + mx - 1, true, slice, g.bd_luma, valid);
    int16_t* sp = &s.y[r * LP + 8 * e];
    const uint4 raw = *reinterpret_cast<const uint4*>(sp);
    int v[8] = {(int)(int16_t)(raw.x & 0xFFFF), (int)(int16_t)(raw.x >> 16), (int)(int16_t)(raw.y & 0xFFFF), (int)(int16_t)(raw.y >> 16),
                (int)(int16_t)(raw.z & 0xFFFF), (int)(int16_t)(raw.z >> 16), (int)(int16_t)(raw.w & 0xFFFF), (int)(int16_t)(raw.w >> 16)};
    if (luma_line(v, ep, e, e + 24, max_y)) {
      uint4 o;
      o.x = (uint32_t)(uint16_t)v[0] | ((uint32_t)(uint16_t)v[1] << 16);
      o.y = (uint32_t)(uint16_t)v[2] | ((uint32_t)(uint16_t)v[3] << 16);
      o.z = (uint32_t)(uint16_t)v[4] | ((uint32_t)(uint16_t)v[5] << 16);
      o.w = (uint32_t)(uint16_t)v[6] | ((uint32_t)(uint16_t)v[7] << 16);
      *reinterpret_cast<uint4*>(sp) = o;
    }
  }
  // ---- vertical edges, chroma: 2 planes x 16 rows x 4 edges = 128 line tasks ----
  if (tid < 128) {
    const int pl = tid >> 6, r = (tid & 63) >> 2, k = tid & 3;
    const int xg = cx0 + 4 + 8 * k, yg = cy0 + r;  // first Q sample (chroma coordinates)
    // chroma row yg covers luma rows 2yg, 2yg+1 -> unit row (2 * yg) >> 2 = yg >> 1; window row = (yg >> 1) - uy0.
    const int mx = 2 + 4 * k, wy = (yg >> 1) - uy0;
    if (xg > 0 && xg < cw && yg >= 0 && yg < crow) {
      const int q = wy * MW + mx;
      const int slice = sd.ctu_slice ? sd.ctu_slice[(size_t)((2 * yg + ctu_row0) >> g.ctu_log2) * g.ctus_w + ((2 * xg) >> g.ctu_log2)] : 0;
      bool no_p, no_q;
      const int tc = chroma_tc(cinfo, prm, q, q - 1, true, slice, pl, g.bd_chroma, no_p, no_q);
      if (tc >= 0) {
        int16_t* sp = &s.c[pl][r * CP + 2 + 8 * k];
        const int m2 = sp[0], m3 = sp[1], m4 = sp[2], m5 = sp[3];
        const int delta = clip3i(-tc, tc, (((m4 - m3) << 2) + m2 - m5 + 4) >> 3);
        if (!no_p) sp[1] = (int16_t)clip3i(0, max_c, m3 + delta);
        if (!no_q) sp[2] = (int16_t)clip3i(0, max_c, m4 - delta);
      }
    }
  }
  __syncthreads();

  // ---- horizontal edges, luma: warp pair per edge; lane = column, 8 segments of 4 columns per warp ----
  {
    const int h = warp >> 1, col = (warp & 1) * 32 + lane;
    const int xg = x0 + col, yg = y0 + 4 + 8 * h;  // first Q sample
    const int mx = 1 + (col >> 2), my = 1 + 2 * h;
    const bool valid = xg >= 0 && xg < g.width && yg > 0 && yg < rows;
    const int slice = (valid && sd.ctu_slice) ? sd.ctu_slice[(size_t)((yg + ctu_row0) >> g.ctu_log2) * g.ctus_w + (xg >> g.ctu_log2)] : 0;
    const EdgeParams ep = luma_edge_params<MV>(s, prm, my * MW + mx, (my - 1) * MW + mx, false, slice, g.bd_luma, valid);
    int16_t* sp = &s.y[(8 * h) * LP + col];
    int v[8];
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = sp[i * LP];
    if (luma_line(v, ep, lane & ~3, lane | 3, max_y)) {
#pragma unroll
      for (int i = 1; i < 7; i++) sp[i * LP] = (int16_t)v[i];
    }
  }
  // ---- horizontal edges, chroma: 2 planes x 2 edges x 32 columns = 128 column tasks ----
  if (tid < 128) {
    const int pl = tid >> 6, h = (tid & 63) >> 5, col = tid & 31;
    const int xg = cx0 + col, yg = cy0 + 2 + 8 * h;  // first Q sample
    if (xg >= 0 && xg < cw && yg > 0 && yg < crow) {
      const int q = ((yg >> 1) - uy0) * MW + ((xg >> 1) - ux0);
      const int slice = sd.ctu_slice ? sd.ctu_slice[(size_t)((2 * yg + ctu_row0) >> g.ctu_log2) * g.ctus_w + ((2 * xg) >> g.ctu_log2)] : 0;
      bool no_p, no_q;
      const int tc = chroma_tc(cinfo, prm, q, q - MW, false, slice, pl, g.bd_chroma, no_p, no_q);
      if (tc >= 0) {
        int16_t* sp = &s.c[pl][(8 * h) * CP + col];
        const int m2 = sp[0], m3 = sp[CP], m4 = sp[2 * CP], m5 = sp[3 * CP];
        const int delta = clip3i(-tc, tc, (((m4 - m3) << 2) + m2 - m5 + 4) >> 3);
        if (!no_p) sp[CP] = (int16_t)clip3i(0, max_c, m3 + delta);
        if (!no_q) sp[2 * CP] = (int16_t)clip3i(0, max_c, m4 - delta);
      }
    }
  }
  __syncthreads();

  // ---- write back ----
#pragma unroll
  for (int i = 0; i < (TH * TW / 4) / NTHREADS; i++) {
    const int c = tid + i * NTHREADS, r = c >> 4, k = c & 15;
    const int x = x0 + k * 4, y = y0 + r;
    if (x >= 0 && x < g.width && y >= 0 && y < rows)
      *reinterpret_cast<uint2*>(dst_y + (size_t)y * g.pitch_y + x) = *reinterpret_cast<const uint2*>(&s.y[r * LP + k * 4]);
  }
  {
    const int pl = tid >> 7, c = tid & 127, r = c >> 3, k = c & 7;
    const int x = cx0 + k * 4, y = cy0 + r;
    if (x >= 0 && x < cw && y >= 0 && y < crow)
      *reinterpret_cast<uint2*>(sd.buf[dst_b][1 + pl] + (size_t)y * g.pitch_c + x) =
          *reinterpret_cast<const uint2*>(&s.c[pl][r * CP + k * 4]);
  }
}

}  // namespace

void launch_deblock(const Geom& g, const SlotDev* slots, int first_slot, int num_slots, int src_b, int dst_b, int mv_mode, cudaStream_t st) {
  dim3 grid((g.width + 4 + TW - 1) / TW, (g.rows + 4 + TH - 1) / TH, num_slots);
  if (mv_mode == 0) deblock_kernel<0><<<grid, NTHREADS, 0, st>>>(g, slots, first_slot, src_b, dst_b);
  else if (mv_mode == 1) deblock_kernel<1><<<grid, NTHREADS, 0, st>>>(g, slots, first_slot, src_b, dst_b);
  else deblock_kernel<2><<<grid, NTHREADS, 0, st>>>(g, slots, first_slot, src_b, dst_b);
}

}  // namespace ilf
