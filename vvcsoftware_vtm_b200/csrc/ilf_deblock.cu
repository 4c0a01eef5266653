// ilf_deblock.cu -- deblocking filter, both edge directions fused in one pass (sm_100a).
//
// Replaces LoopFilter::loopFilterPic (LoopFilter.cpp:149-230): "all vertical edges of the picture, then all
// horizontal edges".  Filtered edges lie on the 8x8 luma grid (:313-324), each reads 4 and writes 3 samples
// per side (:856-916), and the on/off + strong/weak decisions of a 4-line segment use lines 0 and 3 of that
// segment only (:640-671).  Therefore the dependency closure of every 8x8 block SHIFTED by (-4,-4) is the
// block itself: vertical-edge filtering of its 8 rows needs only its 8 columns, and the horizontal edge in
// its middle needs only those 8 vertically filtered rows.  A CTA therefore stages one shifted tile
// (128x32 luma + two 64x16 chroma tiles) in shared memory, runs the vertical pass and then the horizontal
// pass on it, and writes it out: every sample is read from HBM once and written once, with no halo.
// Chroma edges lie on the 8x8 chroma grid and reach 2/1 samples, so any shift in [2,6] closes them; the
// chroma tile is shifted by (-4,-2) to keep 8-byte alignment of its rows.
//
// Per-edge derivation on the device (xGetBoundaryStrengthSingle :419-541, QP/tc/beta :626-634, chroma QP
// :811-829) from the packed per-4x4 grid described in include/ilf_b200.h, staged as a 33x8 unit window.
//
// Work split: the unit of deblocking work is a SEGMENT (4 lines of one edge).  The 128 threads of a CTA take one
// segment each in four phases -- luma vertical (16 edge columns x 8), chroma vertical (2 planes x 8 x 8 units),
// luma horizontal (4 edge rows x 32), chroma horizontal (2 x 2 x 32) -- so edge flag, bS, QP, tc and beta are
// derived once per segment, segments without an edge cost a few instructions, and no lane idles by construction.
#include "ilf_common.cuh"

namespace ilf {
namespace {

constexpr int TW = 128, TH = 32;        // luma tile, origin shifted by (-4, -4)
constexpr int LP = TW + 8;              // luma smem pitch (samples); 272-byte rows keep 16-byte alignment
constexpr int CTW = 64, CTH = 16;       // chroma tile per plane, origin shifted by (-4, -2)
constexpr int CP = CTW + 8;             // chroma smem pitch
constexpr int MW = 33, MH = 8;          // staged metadata window in 4x4 units, origin (32 tx - 2, 8 ty - 1)
constexpr int NTHREADS = 128;           // one task per thread in each of the four phases

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

template <int MV> struct MvStore { uint32_t dummy; };
template <> struct MvStore<1> { uint2 v[MH * MW]; };   // int16 x 4 per unit
template <> struct MvStore<2> { int4 v[MH * MW]; };    // int32 x 4 per unit

template <int MV>
struct Smem {
  int16_t y[TH * LP];
  int16_t c[2][CTH * CP];
  uint32_t info[MH * MW];
  uint32_t info_c[MH * MW];
  MvStore<MV> mv;
};

template <int MV> __device__ __forceinline__ void mv_get(const Smem<MV>& s, int u, int m[4]);
template <> __device__ __forceinline__ void mv_get<0>(const Smem<0>&, int, int m[4]) { m[0] = m[1] = m[2] = m[3] = 0; }
template <> __device__ __forceinline__ void mv_get<1>(const Smem<1>& s, int u, int m[4]) {
  const uint2 v = s.mv.v[u];
  m[0] = (int)(int16_t)(v.x & 0xFFFF); m[1] = (int)(int16_t)(v.x >> 16); m[2] = (int)(int16_t)(v.y & 0xFFFF); m[3] = (int)(int16_t)(v.y >> 16);
}
template <> __device__ __forceinline__ void mv_get<2>(const Smem<2>& s, int u, int m[4]) {
  const int4 v = s.mv.v[u];
  m[0] = v.x; m[1] = v.y; m[2] = v.z; m[3] = v.w;
}

// bS of one 4-sample segment (xGetBoundaryStrengthSingle, LoopFilter.cpp:419-541).  q, p index the staged window.
template <int MV>
__device__ __forceinline__ int boundary_strength(const Smem<MV>& s, uint32_t iq, uint32_t ip, int q, int p, uint32_t tu_bit, int thr) {
  if ((iq | ip) & ILF_BI_INTRA) return 2;
  if ((iq & tu_bit) && ((iq | ip) & ILF_BI_CBF)) return 1;
  const int rq0 = (iq >> 16) & 0xFF, rq1 = iq >> 24, rp0 = (ip >> 16) & 0xFF, rp1 = ip >> 24;
  int mq[4], mp[4];
  mv_get<MV>(s, q, mq);
  mv_get<MV>(s, p, mp);
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

// Luma filter of one line across an edge (xPelFilterLuma, LoopFilter.cpp:856-916).  v[0..7] = p3 p2 p1 p0 | q0 q1 q2 q3.
__device__ __forceinline__ void filter_luma_line(int v[8], int tc, bool sw, bool no_p, bool no_q, int thr_cut, bool second_p, bool second_q, int max_val) {
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

struct EdgeParams {
  int bs;        // 0 = leave the segment alone
  int tc, beta;
  bool no_p, no_q;
};

__device__ __forceinline__ int slice_of(const Geom& g, const SlotDev& sd, int xg, int yg_local) {
  return sd.ctu_slice ? (int)__ldg(sd.ctu_slice + (size_t)((yg_local + g.row0) >> g.ctu_log2) * g.ctus_w + (xg >> g.ctu_log2)) : 0;
}

// Parameters of one luma segment (xEdgeFilterLuma, LoopFilter.cpp:543-634): edge flag, bS, QP, tc, beta.  (xg, yg) = first Q sample.
template <int MV>
__device__ __forceinline__ EdgeParams luma_edge_params(const Smem<MV>& s, const Geom& g, const SlotDev& sd, int q, int p, bool vertical, int xg, int yg) {
  EdgeParams ep;
  ep.bs = 0; ep.tc = 0; ep.beta = 0; ep.no_p = ep.no_q = false;
  const uint32_t iq = s.info[q], ip = s.info[p];
  if (!(iq & (vertical ? ILF_BI_EDGE_V : ILF_BI_EDGE_H))) return ep;
  const ilf_deblock_params* __restrict__ prm = sd.db_params;
  const int bs = boundary_strength<MV>(s, iq, ip, q, p, vertical ? ILF_BI_TU_V : ILF_BI_TU_H, prm->mv_threshold);
  if (!bs) return ep;
  const int slice = slice_of(g, sd, xg, yg);
  const int qp = ((int)(int8_t)(ip >> 8) + (int)(int8_t)(iq >> 8) + 1) >> 1;
  const int tc_off = prm->slices[slice].tc_offset_div2, beta_off = prm->slices[slice].beta_offset_div2;
  const int scale = 1 << (g.bd_luma - 8);
  ep.bs = bs;
  ep.tc = c_tc[clip3i(0, 65, qp + 2 * (bs - 1) + 2 * tc_off)] * scale;
  ep.beta = c_beta[clip3i(0, 63, qp + 2 * beta_off)] * scale;
  ep.no_p = (ip & ILF_BI_NOFILT) != 0;
  ep.no_q = (iq & ILF_BI_NOFILT) != 0;
  return ep;
}

// Decisions of a segment from its lines 0 and 3 (LoopFilter.cpp:640-671) and filtering of its four lines.
// L[i][0..7] = line i of the segment, p3 p2 p1 p0 | q0 q1 q2 q3.  Returns false when nothing was changed.
__device__ __forceinline__ bool filter_luma_segment(int (&L)[4][8], const EdgeParams& ep, int max_val) {
  const int dp0 = abs(L[0][1] - 2 * L[0][2] + L[0][3]), dq0 = abs(L[0][4] - 2 * L[0][5] + L[0][6]);
  const int dp3 = abs(L[3][1] - 2 * L[3][2] + L[3][3]), dq3 = abs(L[3][4] - 2 * L[3][5] + L[3][6]);
  const int d0 = dp0 + dq0, d3 = dp3 + dq3;
  if (d0 + d3 >= ep.beta) return false;
  const int side_thr = (ep.beta + (ep.beta >> 1)) >> 3;
  const bool second_p = (dp0 + dp3) < side_thr, second_q = (dq0 + dq3) < side_thr;
  // xUseStrongFiltering on lines 0 and 3 with d = 2 * d0 / 2 * d3 (:670-671, :960-970)
  const int tc52 = (ep.tc * 5 + 1) >> 1;
  const bool s0 = (abs(L[0][0] - L[0][3]) + abs(L[0][7] - L[0][4]) < (ep.beta >> 3)) && (2 * d0 < (ep.beta >> 2)) && (abs(L[0][3] - L[0][4]) < tc52);
  const bool s3 = (abs(L[3][0] - L[3][3]) + abs(L[3][7] - L[3][4]) < (ep.beta >> 3)) && (2 * d3 < (ep.beta >> 2)) && (abs(L[3][3] - L[3][4]) < tc52);
  const bool sw = s0 && s3;
#pragma unroll
  for (int i = 0; i < 4; i++) filter_luma_line(L[i], ep.tc, sw, ep.no_p, ep.no_q, ep.tc * 10, second_p, second_q, max_val);
  return true;
}

// tc of a chroma segment for one component, or -1 when the segment is not filtered (xEdgeFilterChroma, LoopFilter.cpp:684-838).
__device__ __forceinline__ int chroma_tc(const uint32_t* info, const Geom& g, const SlotDev& sd, int q, int p, bool vertical, int comp, int xg_luma, int yg_luma,
                                         bool& no_p, bool& no_q) {
  const uint32_t iq = info[q], ip = info[p];
  if (!(iq & (vertical ? ILF_BI_EDGE_V : ILF_BI_EDGE_H))) return -1;
  if (!((iq | ip) & ILF_BI_INTRA)) return -1;  // chroma is filtered for bS == 2 only (:769)
  const ilf_deblock_params* __restrict__ prm = sd.db_params;
  const int slice = slice_of(g, sd, xg_luma, yg_luma);
  int qp = (((int)(int8_t)(ip >> 8) + (int)(int8_t)(iq >> 8) + 1) >> 1) + (comp == 0 ? prm->cb_qp_offset : prm->cr_qp_offset);
  if (qp >= 70) qp -= 6;
  else if (qp >= 0) qp = c_chroma_scale[qp];
  no_p = (ip & ILF_BI_NOFILT) != 0;
  no_q = (iq & ILF_BI_NOFILT) != 0;
  return c_tc[clip3i(0, 65, qp + 2 + 2 * prm->slices[slice].tc_offset_div2)] * (1 << (g.bd_chroma - 8));
}

__device__ __forceinline__ void unpack2(uint32_t w, int& a, int& b) { a = (int)(int16_t)(w & 0xFFFF); b = (int)(int16_t)(w >> 16); }
__device__ __forceinline__ uint32_t pack2(int a, int b) { return (uint32_t)(uint16_t)a | ((uint32_t)b << 16); }

template <int MV>
__global__ void __launch_bounds__(NTHREADS) deblock_kernel(Geom g, const SlotDev* __restrict__ slots, int first_slot, BatchCtl bc) {
  __shared__ __align__(16) Smem<MV> s;
  const SlotDev& sd = slots[first_slot + bc.slot[blockIdx.z]];
  const unsigned ctl = bc.v[blockIdx.z];
  const int src_b = ctl_src(ctl, 0), dst_b = ctl_dst(ctl, 0);  // deblocking starts from the uploaded picture: all planes in one buffer
  const int tid = threadIdx.x;
  const int tx = blockIdx.x, ty = blockIdx.y;
  const int x0 = tx * TW - 4, y0 = ty * TH - 4;      // luma tile origin; rows are LOCAL (held-region) rows
  const int cx0 = tx * CTW - 4, cy0 = ty * CTH - 2;  // chroma tile origin
  const int ux0 = tx * (TW / 4) - 2, uy0 = ty * (TH / 4) - 1;  // metadata window origin (units)
  const int rows = g.rows, crow = g.rows >> 1, cw = g.width >> 1;
  const int units_h_local = rows >> 2;
  const bool interior = x0 >= 0 && x0 + TW <= g.width && y0 >= 0 && y0 + TH <= rows;  // then the chroma tile and the unit window are inside too

  // ---- stage: all global loads first (8-byte chunks: the -4 shift keeps 8-byte alignment), then the smem stores ----
  {
    const int16_t* __restrict__ src_y = sd.buf[src_b][0];
    uint2 ly[8], lc[4];
    const int k = tid & 31, r0 = tid >> 5;  // chunk column, first row; rows r0, r0 + 4, ...
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const int r = r0 + 4 * i, x = x0 + k * 4, y = y0 + r;
      ly[i] = make_uint2(0u, 0u);
      if (interior || (x >= 0 && x < g.width && y >= 0 && y < rows)) ly[i] = ldg_u2(src_y + (size_t)y * g.pitch_y + x);
    }
    const int kc = tid & 15, rc0 = tid >> 4;  // chroma: 16 chunks per row, rows rc0 and rc0 + 8, both planes
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int pl = i >> 1, r = rc0 + 8 * (i & 1), x = cx0 + kc * 4, y = cy0 + r;
      lc[i] = make_uint2(0u, 0u);
      if (interior || (x >= 0 && x < cw && y >= 0 && y < crow)) lc[i] = ldg_u2(sd.buf[src_b][1 + pl] + (size_t)y * g.pitch_c + x);
    }
    // metadata window: 264 units
    const bool has_ctree = sd.info_c != nullptr;
#pragma unroll
    for (int i = 0; i < 3; i++) {
      const int m = tid + i * NTHREADS;
      if (m < MH * MW) {
        const int mx = m % MW, my = m / MW, ux = ux0 + mx, uy = uy0 + my;
        const bool in = ux >= 0 && ux < g.units_w && uy >= 0 && uy < units_h_local;
        const size_t u = (size_t)uy * g.units_w + ux;
        s.info[m] = in ? __ldg(sd.info + u) : 0u;
        if (has_ctree) s.info_c[m] = in ? __ldg(sd.info_c + u) : 0u;
        if (MV == 1) reinterpret_cast<uint2*>(&s.mv)[m] = in ? ldg_u2(sd.mv16 + u * 4) : make_uint2(0u, 0u);
        if (MV == 2) reinterpret_cast<uint4*>(&s.mv)[m] = in ? ldg_u4(sd.mv32 + u * 4) : make_uint4(0u, 0u, 0u, 0u);
      }
    }
#pragma unroll
    for (int i = 0; i < 8; i++) *reinterpret_cast<uint2*>(&s.y[(r0 + 4 * i) * LP + k * 4]) = ly[i];
#pragma unroll
    for (int i = 0; i < 4; i++) *reinterpret_cast<uint2*>(&s.c[i >> 1][(rc0 + 8 * (i & 1)) * CP + kc * 4]) = lc[i];
  }
  __syncthreads();

  const uint32_t* cinfo = sd.info_c != nullptr ? s.info_c : s.info;
  const int max_y = (1 << g.bd_luma) - 1, max_c = (1 << g.bd_chroma) - 1;

  // ---- vertical edges, luma: task = 4 lines x 8 samples.  16 edge columns x 8 segment rows ----
  if (!(g.debug & 1)) {
    const int e = tid & 15, sg = tid >> 4;
    const int q = sg * MW + 2 * e + 2;                 // Q unit in the window (x = 128 tx + 8 e -> unit 32 tx + 2 e)
    const int xg = x0 + 4 + 8 * e, yg = y0 + 4 * sg;   // first Q sample
    const EdgeParams ep = luma_edge_params<MV>(s, g, sd, q, q - 1, true, xg, yg);
    if (ep.bs) {
      int16_t* sp = &s.y[(4 * sg) * LP + 8 * e];
      int L[4][8];
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const uint4 raw = *reinterpret_cast<const uint4*>(sp + i * LP);
        unpack2(raw.x, L[i][0], L[i][1]); unpack2(raw.y, L[i][2], L[i][3]); unpack2(raw.z, L[i][4], L[i][5]); unpack2(raw.w, L[i][6], L[i][7]);
      }
      if (filter_luma_segment(L, ep, max_y)) {
#pragma unroll
        for (int i = 0; i < 4; i++)
          *reinterpret_cast<uint4*>(sp + i * LP) = make_uint4(pack2(L[i][0], L[i][1]), pack2(L[i][2], L[i][3]), pack2(L[i][4], L[i][5]), pack2(L[i][6], L[i][7]));
      }
    }
  }
  // ---- vertical edges, chroma: task = one unit = 2 lines x 4 samples.  2 planes x 8 edge columns x 8 unit rows ----
  if (!(g.debug & 1)) {
    const int pl = tid >> 6, k = (tid >> 3) & 7, sg = tid & 7;
    const int q = sg * MW + 4 * k + 2;               // chroma x = 64 tx + 8 k -> luma 128 tx + 16 k -> unit 32 tx + 4 k
    bool no_p, no_q;
    const int tc = chroma_tc(cinfo, g, sd, q, q - 1, true, pl, 2 * (cx0 + 4 + 8 * k), 2 * (cy0 + 2 * sg), no_p, no_q);
    if (tc >= 0) {
#pragma unroll
      for (int i = 0; i < 2; i++) {
        int16_t* sp = &s.c[pl][(2 * sg + i) * CP + 2 + 8 * k];
        const int m2 = sp[0], m3 = sp[1], m4 = sp[2], m5 = sp[3];
        const int delta = clip3i(-tc, tc, (((m4 - m3) << 2) + m2 - m5 + 4) >> 3);
        if (!no_p) sp[1] = (int16_t)clip3i(0, max_c, m3 + delta);
        if (!no_q) sp[2] = (int16_t)clip3i(0, max_c, m4 - delta);
      }
    }
  }
  __syncthreads();

  // ---- horizontal edges, luma: task = 4 columns x 8 rows.  4 edge rows x 32 segments (a warp = one edge row) ----
  if (!(g.debug & 1)) {
    const int sg = tid & 31, h = tid >> 5;
    const int q = (2 * h + 1) * MW + sg + 1;         // unit column 32 tx - 1 + sg, unit row 8 ty + 2 h
    const int xg = x0 + 4 * sg, yg = y0 + 4 + 8 * h;
    const EdgeParams ep = luma_edge_params<MV>(s, g, sd, q, q - MW, false, xg, yg);
    if (ep.bs) {
      int16_t* sp = &s.y[(8 * h) * LP + 4 * sg];
      int L[4][8];
#pragma unroll
      for (int i = 0; i < 8; i++) {
        const uint2 raw = *reinterpret_cast<const uint2*>(sp + i * LP);
        unpack2(raw.x, L[0][i], L[1][i]); unpack2(raw.y, L[2][i], L[3][i]);
      }
      if (filter_luma_segment(L, ep, max_y)) {
#pragma unroll
        for (int i = 1; i < 7; i++) *reinterpret_cast<uint2*>(sp + i * LP) = make_uint2(pack2(L[0][i], L[1][i]), pack2(L[2][i], L[3][i]));
      }
    }
  }
  // ---- horizontal edges, chroma: task = one unit = 2 columns x 4 rows.  2 planes x 2 edge rows x 32 units ----
  if (!(g.debug & 1)) {
    const int pl = tid >> 6, h = (tid >> 5) & 1, m = tid & 31;
    const int q = (4 * h + 1) * MW + m;               // chroma y = 16 ty + 8 h -> luma 32 ty + 16 h -> unit row 8 ty + 4 h
    bool no_p, no_q;
    const int tc = chroma_tc(cinfo, g, sd, q, q - MW, false, pl, 2 * (cx0 + 2 * m), 2 * (cy0 + 2 + 8 * h), no_p, no_q);
    if (tc >= 0) {
      int16_t* sp = &s.c[pl][(8 * h) * CP + 2 * m];
      int a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; i++) unpack2(*reinterpret_cast<const uint32_t*>(sp + i * CP), a[i], b[i]);
      const int da = clip3i(-tc, tc, (((a[2] - a[1]) << 2) + a[0] - a[3] + 4) >> 3);
      const int db = clip3i(-tc, tc, (((b[2] - b[1]) << 2) + b[0] - b[3] + 4) >> 3);
      if (!no_p) *reinterpret_cast<uint32_t*>(sp + CP) = pack2(clip3i(0, max_c, a[1] + da), clip3i(0, max_c, b[1] + db));
      if (!no_q) *reinterpret_cast<uint32_t*>(sp + 2 * CP) = pack2(clip3i(0, max_c, a[2] - da), clip3i(0, max_c, b[2] - db));
    }
  }
  __syncthreads();

  // ---- write back ----
  {
    int16_t* __restrict__ dst_y = sd.buf[dst_b][0];
    const int k = tid & 31, r0 = tid >> 5;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const int r = r0 + 4 * i, x = x0 + k * 4, y = y0 + r;
      if (interior || (x >= 0 && x < g.width && y >= 0 && y < rows))
        *reinterpret_cast<uint2*>(dst_y + (size_t)y * g.pitch_y + x) = *reinterpret_cast<const uint2*>(&s.y[r * LP + k * 4]);
    }
    const int kc = tid & 15, rc0 = tid >> 4;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int pl = i >> 1, r = rc0 + 8 * (i & 1), x = cx0 + kc * 4, y = cy0 + r;
      if (interior || (x >= 0 && x < cw && y >= 0 && y < crow))
        *reinterpret_cast<uint2*>(sd.buf[dst_b][1 + pl] + (size_t)y * g.pitch_c + x) = *reinterpret_cast<const uint2*>(&s.c[pl][r * CP + kc * 4]);
    }
  }
}

}  // namespace

void launch_deblock(const Geom& g, const SlotDev* slots, int first_slot, int num_slots, const BatchCtl& ctl, int mv_mode, cudaStream_t st) {
  dim3 grid((g.width + 4 + TW - 1) / TW, (g.rows + 4 + TH - 1) / TH, num_slots);
  if (mv_mode == 0) deblock_kernel<0><<<grid, NTHREADS, 0, st>>>(g, slots, first_slot, ctl);
  else if (mv_mode == 1) deblock_kernel<1><<<grid, NTHREADS, 0, st>>>(g, slots, first_slot, ctl);
  else deblock_kernel<2><<<grid, NTHREADS, 0, st>>>(g, slots, first_slot, ctl);
}

}  // namespace ilf
