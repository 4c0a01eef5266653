// ilf_alf.cu -- adaptive loop filter: 4x4 block classification + 7x7/5x5 luma and 5x5 chroma diamond filters (sm_100a).
//
// Replaces AdaptiveLoopFilter::ALFProcess after coefficient reconstruction (AdaptiveLoopFilter.cpp:89-138):
//   deriveClassificationBlk :292-463   Laplacian activity/direction -> classIdx (0..24), transposeIdx (0..3)
//   filterBlk<7/5>          :465-650   point-symmetric diamond FIR, (sum + 256) >> 9, clip to [0, 2^bd - 1]
// The source is the whole SAO'd picture padded by 3 replicated samples (:90-92, Buffer.h:433-465); here the
// padding is coordinate clamping while a tile + 3-sample halo is staged in shared memory.  CTUs whose enable
// flag is 0 are copied through (the stage reads one buffer and writes the other).
//
// Luma: one CTA (256 threads) per 128x32 tile; a thread owns one 4x4 block from classification to output.
//   phase 1  stage the tile + 3-sample halo in shared memory as int32 (38 rows x 136 columns), all global loads first
//   phase 2  1-D Laplacians of every sample, summed per 2x2 cell (two 16-bit sums per word); a task = 4 rows x 8 columns
//   phase 3  each thread sums the 4x4 cells of its block's 8x8 window and derives class + transpose (kept in a register)
//   phase 4  each thread filters its block with the (class, transpose) coefficient row of the per-picture table that
//            ilf_set_alf_params precomputed (SlotDev::alf_coef), and stores four int16x4 rows
// Tiles whose blocks are all in CTUs with ALF off are copied through without touching shared memory.
#include "ilf_common.cuh"

namespace ilf {
namespace {

constexpr int LT_W = 128, LT_H = 32;         // luma tile
constexpr int LS_W = LT_W + 8;               // staged columns: x0-4 .. x0+131
constexpr int LS_P = LS_W + 4;               // smem pitch in samples (phase 2 reads one aligned int4 past the staged columns)
constexpr int LS_H = LT_H + 6;               // staged rows:    y0-3 .. y0+34
constexpr int CELL_W = LT_W / 2 + 2, CELL_H = LT_H / 2 + 2;  // 66 x 18 cells of 2x2 samples, first cell at (x0-2, y0-2)
constexpr int NT = 256;

__constant__ uint8_t c_th[16] = {0, 1, 2, 2, 2, 2, 2, 3, 3, 3, 3, 3, 3, 3, 3, 4};
__constant__ uint8_t c_transpose[8] = {0, 1, 0, 2, 2, 3, 1, 3};

struct LumaSmem {
  int t[LS_H][LS_P];               // samples as int32
  uint2 cell[CELL_H][CELL_W];      // {V | H << 16, D0 | D1 << 16} per 2x2 cell
};

// Class of one 4x4 block from its four window sums (AdaptiveLoopFilter.cpp:390-451).
__device__ __forceinline__ int classify(int sum_v, int sum_h, int sum_d0, int sum_d1, int shift) {
  const int activity = clip3i(0, 15, ((sum_v + sum_h) * 32) >> shift);
  int class_idx = c_th[activity];
  int hv1, hv0, d1, d0, dir_hv, dir_d;
  if (sum_v > sum_h) { hv1 = sum_v; hv0 = sum_h; dir_hv = 1; } else { hv1 = sum_h; hv0 = sum_v; dir_hv = 3; }
  if (sum_d0 > sum_d1) { d1 = sum_d0; d0 = sum_d1; dir_d = 0; } else { d1 = sum_d1; d0 = sum_d0; dir_d = 2; }
  int hvd1, hvd0, main_dir, sec_dir;
  // the reference multiplies in `int` and x86 wraps mod 2^32 (:420): multiply unsigned, compare signed
  if ((int)((unsigned)d1 * (unsigned)hv0) > (int)((unsigned)hv1 * (unsigned)d0)) {
    hvd1 = d1; hvd0 = d0; main_dir = dir_d; sec_dir = dir_hv;
  } else {
    hvd1 = hv1; hvd0 = hv0; main_dir = dir_hv; sec_dir = dir_d;
  }
  int strength = 0;
  if (hvd1 > 2 * hvd0) strength = 1;
  if (hvd1 * 2 > 9 * hvd0) strength = 2;
  if (strength) class_idx += (((main_dir & 1) << 1) + strength) * 5;
  return class_idx | (c_transpose[main_dir * 2 + (sec_dir >> 1)] << 5);
}

__device__ __forceinline__ void ld12(const int* p, int w[12]) {
  const int4 a = *reinterpret_cast<const int4*>(p), b = *reinterpret_cast<const int4*>(p + 4), c = *reinterpret_cast<const int4*>(p + 8);
  w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w; w[8] = c.x; w[9] = c.y; w[10] = c.z; w[11] = c.w;
}

template <bool CLASSIFY_ONLY>
__global__ void __launch_bounds__(NT) alf_luma_kernel(Geom g, const SlotDev* __restrict__ slots, int first_slot, BatchCtl bc) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  LumaSmem& s = *reinterpret_cast<LumaSmem*>(smem_raw);
  const SlotDev& sd = slots[first_slot + bc.slot[blockIdx.z]];
  const unsigned ctl = bc.v[blockIdx.z];
  if (ctl_skip(ctl, 0)) return;
  const int tid = threadIdx.x;
  const int x0 = blockIdx.x * LT_W, y0 = blockIdx.y * LT_H;  // local rows
  const int16_t* __restrict__ src = sd.buf[ctl_src(ctl, 0)][0];
  int16_t* __restrict__ dst = sd.buf[ctl_dst(ctl, 0)][0];
  const int rows = g.rows;

  // this thread's 4x4 block: a warp = one row of 32 blocks
  const int bj = tid & 31, bi = tid >> 5;
  const int bx = x0 + 4 * bj, by = y0 + 4 * bi;
  const bool blk_in = bx < g.width && by < rows;
  bool en = false;
  if (blk_in) en = CLASSIFY_ONLY ? true : (__ldg(sd.alf_ctu_enable + ((by + g.row0) >> g.ctu_log2) * g.ctus_w + (bx >> g.ctu_log2)) != 0);
  if (!__syncthreads_or(en)) {
    // every CTU under this tile has ALF off: copy through (int16x8 vectors, 2 per thread)
    uint4 v[2];
#pragma unroll
    for (int i = 0; i < 2; i++) {
      const int c = tid + i * NT, r = c >> 4, k = c & 15, x = x0 + 8 * k, y = y0 + r;
      if (x < g.width && y < rows) v[i] = ldg_u4(src + (size_t)y * g.pitch_y + x);
    }
#pragma unroll
    for (int i = 0; i < 2; i++) {
      const int c = tid + i * NT, r = c >> 4, k = c & 15, x = x0 + 8 * k, y = y0 + r;
      if (x < g.width && y < rows) *reinterpret_cast<uint4*>(dst + (size_t)y * g.pitch_y + x) = v[i];
    }
    return;
  }

  // ---- phase 1: stage tile + halo as int32; picture borders replicate (coordinate clamping = extendBorderPel) ----
  {
    constexpr int CHUNKS = LS_H * (LS_W / 4), ROUNDS = (CHUNKS + NT - 1) / NT;
    const bool interior = x0 >= 4 && x0 + LT_W + 4 <= g.width && y0 >= 3 && y0 + LT_H + 3 <= rows;
    uint2 raw[ROUNDS];
#pragma unroll
    for (int i = 0; i < ROUNDS; i++) {
      const int c = tid + i * NT;
      if (c < CHUNKS) {
        const int r = c / (LS_W / 4), k = c % (LS_W / 4);
        int y = y0 - 3 + r, x = x0 - 4 + 4 * k;
        if (interior) raw[i] = ldg_u2(src + (size_t)y * g.pitch_y + x);
        else {
          y = min(max(y, 0), rows - 1);
          const int16_t* rowp = src + (size_t)y * g.pitch_y;
          if (x < 0) { const uint32_t e = (uint16_t)__ldg(rowp); raw[i] = make_uint2(e | (e << 16), e | (e << 16)); }
          else if (x >= g.width) { const uint32_t e = (uint16_t)__ldg(rowp + g.width - 1); raw[i] = make_uint2(e | (e << 16), e | (e << 16)); }
          else raw[i] = ldg_u2(rowp + x);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < ROUNDS; i++) {
      const int c = tid + i * NT;
      if (c < CHUNKS) {
        const int r = c / (LS_W / 4), k = c % (LS_W / 4);
        *reinterpret_cast<int4*>(&s.t[r][4 * k]) = make_int4((int)(int16_t)(raw[i].x & 0xFFFF), (int)(int16_t)(raw[i].x >> 16),
                                                             (int)(int16_t)(raw[i].y & 0xFFFF), (int)(int16_t)(raw[i].y >> 16));
      }
    }
  }
  // coefficient row prefetch does not depend on shared memory; issued after the class is known (phase 4)
  __syncthreads();

  // ---- phase 2: Laplacians per 2x2 cell.  Task = 2 cell rows x 4 cell columns = sample rows 4 tr.., columns 8 tc.. of the
  //      region that starts at (x0-2, y0-2), i.e. staged rows 1 + 4 tr .., staged columns 2 + 8 tc .. ----
  if (tid < (CELL_H / 2) * ((CELL_W + 3) / 4)) {
    constexpr int TPR = (CELL_W + 3) / 4;  // 17 tasks per row, the last one covers 2 cell columns
    const int tr = tid / TPR, tc = tid % TPR;
    const int ncell = tc == TPR - 1 ? CELL_W - 4 * (TPR - 1) : 4;
    int up[12], cur[12], dn[12];
    const int* base = &s.t[4 * tr][8 * tc];
    ld12(base, up);
    ld12(base + LS_P, cur);
    uint32_t acc_vh[4], acc_d[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {  // sample row i of the task; window index j + 2 <-> sample column j
      ld12(base + (i + 2) * LS_P, dn);
      if ((i & 1) == 0) {
#pragma unroll
        for (int k = 0; k < 4; k++) { acc_vh[k] = 0; acc_d[k] = 0; }
      }
#pragma unroll
      for (int j = 0; j < 8; j++) {
        const int c2 = cur[j + 2] << 1;
        const int v = abs(c2 - up[j + 2] - dn[j + 2]);
        const int h = abs(c2 - cur[j + 1] - cur[j + 3]);
        const int d0 = abs(c2 - up[j + 1] - dn[j + 3]);
        const int d1 = abs(c2 - up[j + 3] - dn[j + 1]);
        acc_vh[j >> 1] += (uint32_t)v + ((uint32_t)h << 16);
        acc_d[j >> 1] += (uint32_t)d0 + ((uint32_t)d1 << 16);
      }
      if (i & 1) {
        const int cr = 2 * tr + (i >> 1);
#pragma unroll
        for (int k = 0; k < 4; k++)
          if (k < ncell) s.cell[cr][4 * tc + k] = make_uint2(acc_vh[k], acc_d[k]);
      }
#pragma unroll
      for (int k = 0; k < 12; k++) { up[k] = cur[k]; cur[k] = dn[k]; }
    }
  }
  __syncthreads();

  // ---- phase 3: 4x4 cells of the block's 8x8 window -> class.  A 2x2 cell sum is < 2^15 up to 12 bit, so two cells add
  //      without a carry between the 16-bit halves; wider sums are taken in 32 bits ----
  int sv = 0, sh = 0, sd0 = 0, sd1 = 0;
#pragma unroll
  for (int r = 0; r < 4; r++) {
    const uint4 c01 = *reinterpret_cast<const uint4*>(&s.cell[2 * bi + r][2 * bj]);      // cells 0, 1: {vh0, d0, vh1, d1}
    const uint4 c23 = *reinterpret_cast<const uint4*>(&s.cell[2 * bi + r][2 * bj + 2]);
    const uint32_t a0 = c01.x + c01.z, a1 = c23.x + c23.z, b0 = c01.y + c01.w, b1 = c23.y + c23.w;
    sv += (a0 & 0xFFFF) + (a1 & 0xFFFF); sh += (a0 >> 16) + (a1 >> 16);
    sd0 += (b0 & 0xFFFF) + (b1 & 0xFFFF); sd1 += (b0 >> 16) + (b1 >> 16);
  }
  const int cl = classify(sv, sh, sd0, sd1, g.bd_luma + 4);
  if (CLASSIFY_ONLY) {
    const int ux = (x0 >> 2) + bj, uy = (y0 >> 2) + bi;
    if (ux < g.units_w && uy < (rows >> 2)) sd.alf_class[(size_t)uy * g.units_w + ux] = (uint8_t)cl;
    return;
  }

  // ---- phase 4: filter the block ----
  if (!blk_in) return;
  const int max_val = (1 << g.bd_luma) - 1;
  const int tcx = 4 + 4 * bj;  // staged column of the block's first sample
  if (!en) {
#pragma unroll
    for (int rr = 0; rr < 4; rr++) {
      const int4 v = *reinterpret_cast<const int4*>(&s.t[3 + 4 * bi + rr][tcx]);
      *reinterpret_cast<uint2*>(dst + (size_t)(by + rr) * g.pitch_y + bx) = make_uint2((uint32_t)v.x | ((uint32_t)v.y << 16), (uint32_t)v.z | ((uint32_t)v.w << 16));
    }
    return;
  }
  int f[16];
  {
    const int4* cp = reinterpret_cast<const int4*>(sd.alf_coef + ((cl & 31) * 4 + (cl >> 5)) * 16);
    const int4 a = __ldg(cp), b = __ldg(cp + 1), c = __ldg(cp + 2), d = __ldg(cp + 3);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
    f[8] = c.x; f[9] = c.y; f[10] = c.z; f[11] = c.w; f[12] = d.x; f[13] = d.y; f[14] = d.z; f[15] = d.w;
  }
  const bool is7 = sd.alf->luma_filter_7x7 != 0;
#pragma unroll
  for (int rr = 0; rr < 4; rr++) {
    const int trow = 3 + 4 * bi + rr;  // staged row of the output row
    // rX[i] = staged columns tcx-4 .. tcx+7 (i = 0..11); output sample j reads i = j + 4 + dx
    int sum[4], r0[12], rp[12], rm[12];
    ld12(&s.t[trow][tcx - 4], r0);
    if (is7) {
#pragma unroll
      for (int j = 0; j < 4; j++)
        sum[j] = f[12] * r0[j + 4] + f[11] * (r0[j + 5] + r0[j + 3]) + f[10] * (r0[j + 6] + r0[j + 2]) + f[9] * (r0[j + 7] + r0[j + 1]);
      ld12(&s.t[trow + 1][tcx - 4], rp); ld12(&s.t[trow - 1][tcx - 4], rm);
#pragma unroll
      for (int j = 0; j < 4; j++)
        sum[j] += f[4] * (rp[j + 6] + rm[j + 2]) + f[5] * (rp[j + 5] + rm[j + 3]) + f[6] * (rp[j + 4] + rm[j + 4]) +
                  f[7] * (rp[j + 3] + rm[j + 5]) + f[8] * (rp[j + 2] + rm[j + 6]);
      ld12(&s.t[trow + 2][tcx - 4], rp); ld12(&s.t[trow - 2][tcx - 4], rm);
#pragma unroll
      for (int j = 0; j < 4; j++)
        sum[j] += f[1] * (rp[j + 5] + rm[j + 3]) + f[2] * (rp[j + 4] + rm[j + 4]) + f[3] * (rp[j + 3] + rm[j + 5]);
      {
        const int4 a = *reinterpret_cast<const int4*>(&s.t[trow + 3][tcx]);
        const int4 b = *reinterpret_cast<const int4*>(&s.t[trow - 3][tcx]);
        sum[0] += f[0] * (a.x + b.x); sum[1] += f[0] * (a.y + b.y); sum[2] += f[0] * (a.z + b.z); sum[3] += f[0] * (a.w + b.w);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; j++) sum[j] = f[6] * r0[j + 4] + f[5] * (r0[j + 5] + r0[j + 3]) + f[4] * (r0[j + 6] + r0[j + 2]);
      ld12(&s.t[trow + 1][tcx - 4], rp); ld12(&s.t[trow - 1][tcx - 4], rm);
#pragma unroll
      for (int j = 0; j < 4; j++) sum[j] += f[1] * (rp[j + 5] + rm[j + 3]) + f[2] * (rp[j + 4] + rm[j + 4]) + f[3] * (rp[j + 3] + rm[j + 5]);
      {
        const int4 a = *reinterpret_cast<const int4*>(&s.t[trow + 2][tcx]);
        const int4 b = *reinterpret_cast<const int4*>(&s.t[trow - 2][tcx]);
        sum[0] += f[0] * (a.x + b.x); sum[1] += f[0] * (a.y + b.y); sum[2] += f[0] * (a.z + b.z); sum[3] += f[0] * (a.w + b.w);
      }
    }
    int o4[4];
#pragma unroll
    for (int j = 0; j < 4; j++) o4[j] = clip3i(0, max_val, (sum[j] + 256) >> 9);
    *reinterpret_cast<uint2*>(dst + (size_t)(by + rr) * g.pitch_y + bx) = make_uint2((uint32_t)o4[0] | ((uint32_t)o4[1] << 16), (uint32_t)o4[2] | ((uint32_t)o4[3] << 16));
  }
}

// ---------------------------------------------------------------------------------------------------------
// Chroma: 5x5 diamond, one filter per picture, no classification.  Tile 64x16 per plane, 4 samples per thread.
// ---------------------------------------------------------------------------------------------------------
constexpr int CT_W = 64, CT_H = 16, CS_W = CT_W + 8, CS_H = CT_H + 4;

__global__ void __launch_bounds__(NT) alf_chroma_kernel(Geom g, const SlotDev* __restrict__ slots, int first_slot, BatchCtl bc) {
  __shared__ __align__(16) int t[CS_H][CS_W];
  const SlotDev& sd = slots[first_slot + bc.slot[blockIdx.z >> 1]];
  const int plane = 1 + (blockIdx.z & 1);
  const unsigned ctl = bc.v[blockIdx.z >> 1];
  if (ctl_skip(ctl, plane)) return;
  const int tid = threadIdx.x;
  const int x0 = blockIdx.x * CT_W, y0 = blockIdx.y * CT_H;
  const int cw = g.width >> 1, crows = g.rows >> 1;
  const int16_t* __restrict__ src = sd.buf[ctl_src(ctl, plane)][plane];
  int16_t* __restrict__ dst = sd.buf[ctl_dst(ctl, plane)][plane];
  const int r = tid >> 4, k = tid & 15;  // output: row r, columns 4k..4k+3 of the tile
  const int x = x0 + 4 * k, y = y0 + r;
  const bool in = x < cw && y < crows;
  bool en = false;
  if (in) {
    const int ctu = ((((y << 1) + g.row0) >> g.ctu_log2) * g.ctus_w) + ((x << 1) >> g.ctu_log2);
    en = sd.alf_ctu_enable[(size_t)plane * g.ctus_w * g.ctus_h + ctu] != 0;
  }
  const int any_en = __syncthreads_or(en);
  if (!any_en) {
    if (in) *reinterpret_cast<uint2*>(dst + (size_t)y * g.pitch_c + x) = ldg_u2(src + (size_t)y * g.pitch_c + x);
    return;
  }
  for (int c = tid; c < CS_H * (CS_W / 4); c += NT) {
    const int rr = c / (CS_W / 4), kk = c % (CS_W / 4);
    const int yy = min(max(y0 - 2 + rr, 0), crows - 1);
    const int xx = x0 - 4 + 4 * kk;
    const int16_t* rowp = src + (size_t)yy * g.pitch_c;
    int4 v;
    if (xx < 0) { const int e = rowp[0]; v = make_int4(e, e, e, e); }
    else if (xx >= cw) { const int e = rowp[cw - 1]; v = make_int4(e, e, e, e); }
    else {
      const uint2 raw = ldg_u2(rowp + xx);
      v = make_int4((int)(int16_t)(raw.x & 0xFFFF), (int)(int16_t)(raw.x >> 16), (int)(int16_t)(raw.y & 0xFFFF), (int)(int16_t)(raw.y >> 16));
    }
    *reinterpret_cast<int4*>(&t[rr][4 * kk]) = v;
  }
  __syncthreads();
  if (!in) return;
  const int trow = 2 + r, tcx = 4 + 4 * k;
  if (!en) {
    const int4 v = *reinterpret_cast<const int4*>(&t[trow][tcx]);
    uint2 o; o.x = (uint32_t)(uint16_t)v.x | ((uint32_t)(uint16_t)v.y << 16); o.y = (uint32_t)(uint16_t)v.z | ((uint32_t)(uint16_t)v.w << 16);
    *reinterpret_cast<uint2*>(dst + (size_t)y * g.pitch_c + x) = o;
    return;
  }
  int f[7];
#pragma unroll
  for (int i = 0; i < 7; i++) f[i] = sd.alf->chroma_coeff[i];
  auto load12 = [&](int* w, int row) {
    const int4* p = reinterpret_cast<const int4*>(&t[row][tcx - 4]);
    const int4 a = p[0], b = p[1], c = p[2];
    w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w; w[8] = c.x; w[9] = c.y; w[10] = c.z; w[11] = c.w;
  };
  int r0[12], rp[12], rm[12], sum[4];
  load12(r0, trow);
#pragma unroll
  for (int j = 0; j < 4; j++) sum[j] = f[6] * r0[j + 4] + f[5] * (r0[j + 5] + r0[j + 3]) + f[4] * (r0[j + 6] + r0[j + 2]);
  load12(rp, trow + 1); load12(rm, trow - 1);
#pragma unroll
  for (int j = 0; j < 4; j++) sum[j] += f[1] * (rp[j + 5] + rm[j + 3]) + f[2] * (rp[j + 4] + rm[j + 4]) + f[3] * (rp[j + 3] + rm[j + 5]);
  {
    const int4 a = *reinterpret_cast<const int4*>(&t[trow + 2][tcx]);
    const int4 b = *reinterpret_cast<const int4*>(&t[trow - 2][tcx]);
    sum[0] += f[0] * (a.x + b.x); sum[1] += f[0] * (a.y + b.y); sum[2] += f[0] * (a.z + b.z); sum[3] += f[0] * (a.w + b.w);
  }
  const int max_val = (1 << g.bd_chroma) - 1;
  int o4[4];
#pragma unroll
  for (int j = 0; j < 4; j++) o4[j] = clip3i(0, max_val, (sum[j] + 256) >> 9);
  uint2 o; o.x = (uint32_t)o4[0] | ((uint32_t)o4[1] << 16); o.y = (uint32_t)o4[2] | ((uint32_t)o4[3] << 16);
  *reinterpret_cast<uint2*>(dst + (size_t)y * g.pitch_c + x) = o;
}

}  // namespace

void launch_alf_luma(const Geom& g, const SlotDev* slots, int first_slot, int num_slots, const BatchCtl& ctl, bool classify_only, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(alf_luma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(LumaSmem));
    cudaFuncSetAttribute(alf_luma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(LumaSmem));
    attr_set = true;
  }
  dim3 gl((g.width + LT_W - 1) / LT_W, (g.rows + LT_H - 1) / LT_H, num_slots);
  if (classify_only) {
    alf_luma_kernel<true><<<gl, NT, sizeof(LumaSmem), st>>>(g, slots, first_slot, ctl);
    return;
  }
  alf_luma_kernel<false><<<gl, NT, sizeof(LumaSmem), st>>>(g, slots, first_slot, ctl);
}

void launch_alf_chroma(const Geom& g, const SlotDev* slots, int first_slot, int num_slots, const BatchCtl& ctl, cudaStream_t st) {
  dim3 gc((g.width / 2 + CT_W - 1) / CT_W, (g.rows / 2 + CT_H - 1) / CT_H, 2 * num_slots);
  alf_chroma_kernel<<<gc, NT, 0, st>>>(g, slots, first_slot, ctl);
}

}  // namespace ilf
