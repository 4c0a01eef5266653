// ilf_alf.cu -- adaptive loop filter: 4x4 block classification + 7x7/5x5 luma and 5x5 chroma diamond filters (sm_100a).
//
// Replaces AdaptiveLoopFilter::ALFProcess after coefficient reconstruction (AdaptiveLoopFilter.cpp:89-138):
//   deriveClassificationBlk :292-463   Laplacian activity/direction -> classIdx (0..24), transposeIdx (0..3)
//   filterBlk<7/5>          :465-650   point-symmetric diamond FIR, (sum + 256) >> 9, clip to [0, 2^bd - 1]
// The source is the whole SAO'd picture padded by 3 replicated samples (:90-92, Buffer.h:433-465); here the
// padding is coordinate clamping while a tile + 3-sample halo is staged in shared memory.  CTUs whose enable
// flag is 0 are copied through (the stage reads one buffer and writes the other).
//
// Luma: one CTA per 64x32 tile.  Phase 1 stages the tile as int32; phase 2 computes the four 1-D Laplacians of
// every sample per 2x2 cell (two 16-bit sums per word); phase 3 sums 4x4 cells per 4x4 block and derives the
// class; phase 4 filters, two rows of a 4x4 block per thread, coefficients pre-transposed in shared memory.
#include "ilf_common.cuh"

namespace ilf {
namespace {

constexpr int LT_W = 64, LT_H = 32;          // luma tile
constexpr int LS_W = LT_W + 8;               // staged columns: x0-4 .. x0+67
constexpr int LS_H = LT_H + 6;               // staged rows:    y0-3 .. y0+34
constexpr int CELL_W = LT_W / 2 + 2, CELL_H = LT_H / 2 + 2;
constexpr int NT = 256;

// Coefficient order after transposition (AdaptiveLoopFilter.cpp:541-575).
__constant__ uint8_t c_perm7[4][13] = {{0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12},
                                       {9, 4, 10, 8, 1, 5, 11, 7, 3, 0, 2, 6, 12},
                                       {0, 3, 2, 1, 8, 7, 6, 5, 4, 9, 10, 11, 12},
                                       {9, 8, 10, 4, 3, 7, 11, 5, 1, 0, 2, 6, 12}};
__constant__ uint8_t c_perm5[4][7] = {{0, 1, 2, 3, 4, 5, 6}, {4, 1, 5, 3, 0, 2, 6}, {0, 3, 2, 1, 4, 5, 6}, {4, 3, 5, 1, 0, 2, 6}};
__constant__ uint8_t c_th[16] = {0, 1, 2, 2, 2, 2, 2, 3, 3, 3, 3, 3, 3, 3, 3, 4};
__constant__ uint8_t c_transpose[8] = {0, 1, 0, 2, 2, 3, 1, 3};

struct LumaSmem {
  int t[LS_H][LS_W];               // samples as int32
  uint32_t cell_vh[CELL_H][CELL_W];  // V | H << 16 per 2x2 cell
  uint32_t cell_d[CELL_H][CELL_W];   // D0 | D1 << 16
  int coef[25][4][16];             // [class][transpose][tap], transposition already applied
  uint8_t cls[LT_H / 4][LT_W / 4];
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

template <bool CLASSIFY_ONLY>
__global__ void __launch_bounds__(NT) alf_luma_kernel(Geom g, const SlotDev* __restrict__ slots, int first_slot, BatchCtl bc) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  LumaSmem& s = *reinterpret_cast<LumaSmem*>(smem_raw);
  const SlotDev& sd = slots[first_slot + blockIdx.z];
  const unsigned ctl = bc.v[blockIdx.z];
  if (ctl_skip(ctl, 0)) return;
  const int tid = threadIdx.x;
  const int x0 = blockIdx.x * LT_W, y0 = blockIdx.y * LT_H;  // local rows
  const int16_t* __restrict__ src = sd.buf[ctl_src(ctl, 0)][0];
  int16_t* __restrict__ dst = sd.buf[ctl_dst(ctl, 0)][0];
  const int rows = g.rows;

  // filter phase mapping: thread = (4x4 block, upper/lower half)
  const int blk = tid >> 1, half = tid & 1;
  const int bi = blk >> 4, bj = blk & 15;
  const int bx = x0 + 4 * bj, by = y0 + 4 * bi;
  const bool blk_in = bx < g.width && by < rows;
  bool en = false;
  if (blk_in) {
    const int ctu = (((by + g.row0) >> g.ctu_log2) * g.ctus_w) + (bx >> g.ctu_log2);
    en = CLASSIFY_ONLY ? true : (sd.alf_ctu_enable[ctu] != 0);
  }
  const int any_en = __syncthreads_or(en);
  if (!any_en) {
    // every CTU under this tile has ALF off: copy through (int16x8 vectors)
    for (int c = tid; c < LT_H * (LT_W / 8); c += NT) {
      const int r = c >> 3, k = c & 7, x = x0 + 8 * k, y = y0 + r;
      if (x < g.width && y < rows)
        *reinterpret_cast<uint4*>(dst + (size_t)y * g.pitch_y + x) = ldg_u4(src + (size_t)y * g.pitch_y + x);
    }
    return;
  }

  // ---- phase 1: stage tile + halo as int32, coordinates clamped (= replicate padding) ----
  for (int c = tid; c < LS_H * (LS_W / 4); c += NT) {
    const int r = c / (LS_W / 4), k = c % (LS_W / 4);
    const int y = min(max(y0 - 3 + r, 0), rows - 1);
    const int x = x0 - 4 + 4 * k;
    const int16_t* rowp = src + (size_t)y * g.pitch_y;
    int4 v;
    if (x < 0) { const int e = rowp[0]; v = make_int4(e, e, e, e); }
    else if (x >= g.width) { const int e = rowp[g.width - 1]; v = make_int4(e, e, e, e); }
    else {
      const uint2 raw = ldg_u2(rowp + x);
      v = make_int4((int)(int16_t)(raw.x & 0xFFFF), (int)(int16_t)(raw.x >> 16), (int)(int16_t)(raw.y & 0xFFFF), (int)(int16_t)(raw.y >> 16));
    }
    *reinterpret_cast<int4*>(&s.t[r][4 * k]) = v;
  }
  if (!CLASSIFY_ONLY) {
    const ilf_alf_params* __restrict__ ap = sd.alf;
    const bool is7 = ap->luma_filter_7x7 != 0;
    for (int i = tid; i < 25 * 4 * 16; i += NT) {
      const int cl = i >> 6, tr = (i >> 4) & 3, k = i & 15;
      int v = 0;
      if (is7) { if (k < 13) v = ap->luma_coeff[cl][c_perm7[tr][k]]; }
      else     { if (k < 7) v = ap->luma_coeff[cl][c_perm5[tr][k]]; }
      s.coef[cl][tr][k] = v;
    }
  }
  __syncthreads();

  // ---- phase 2: Laplacians per 2x2 cell.  Cell (ci, cj) = samples rows y0-2+2ci.., cols x0-2+2cj.. ----
  for (int c = tid; c < CELL_H * CELL_W; c += NT) {
    const int ci = c / CELL_W, cj = c % CELL_W;
    const int tr = 1 + 2 * ci, tc = 2 + 2 * cj;  // tile coordinates of the cell's first sample
    int v = 0, h = 0, d0 = 0, d1 = 0;
#pragma unroll
    for (int dy = 0; dy < 2; dy++)
#pragma unroll
      for (int dx = 0; dx < 2; dx++) {
        const int r = tr + dy, cc = tc + dx;
        const int p2 = s.t[r][cc] << 1;
        v += abs(p2 - s.t[r - 1][cc] - s.t[r + 1][cc]);
        h += abs(p2 - s.t[r][cc - 1] - s.t[r][cc + 1]);
        d0 += abs(p2 - s.t[r - 1][cc - 1] - s.t[r + 1][cc + 1]);
        d1 += abs(p2 - s.t[r - 1][cc + 1] - s.t[r + 1][cc - 1]);
      }
    s.cell_vh[ci][cj] = (uint32_t)v | ((uint32_t)h << 16);
    s.cell_d[ci][cj] = (uint32_t)d0 | ((uint32_t)d1 << 16);
  }
  __syncthreads();

  // ---- phase 3: 4x4 cells per block -> class ----
  if (tid < (LT_H / 4) * (LT_W / 4)) {
    const int ci = tid >> 4, cj = tid & 15;  // block row / column in the tile
    int sv = 0, sh = 0, sd0 = 0, sd1 = 0;
#pragma unroll
    for (int r = 0; r < 4; r++) {
      uint32_t a = 0, b = 0;  // 4 cells of <= 8184 per half: no carry between the halves
#pragma unroll
      for (int c = 0; c < 4; c++) { a += s.cell_vh[2 * ci + r][2 * cj + c]; b += s.cell_d[2 * ci + r][2 * cj + c]; }
      sv += a & 0xFFFF; sh += a >> 16; sd0 += b & 0xFFFF; sd1 += b >> 16;
    }
    const int cl = classify(sv, sh, sd0, sd1, g.bd_luma + 4);
    s.cls[ci][cj] = (uint8_t)cl;
    if (CLASSIFY_ONLY) {
      const int ux = (x0 >> 2) + cj, uy = (y0 >> 2) + ci;
      if (ux < g.units_w && uy < (rows >> 2)) sd.alf_class[(size_t)uy * g.units_w + ux] = (uint8_t)cl;
    }
  }
  if (CLASSIFY_ONLY) return;
  __syncthreads();

  // ---- phase 4: filter two rows of a 4x4 block per thread ----
  if (!blk_in) return;
  const int max_val = (1 << g.bd_luma) - 1;
  const int tcx = 4 + 4 * bj;  // tile column of the block's first sample
  if (!en) {
#pragma unroll
    for (int rr = 0; rr < 2; rr++) {
      const int trow = 3 + 4 * bi + 2 * half + rr;
      const int4 v = *reinterpret_cast<const int4*>(&s.t[trow][tcx]);
      uint2 o; o.x = (uint32_t)(uint16_t)v.x | ((uint32_t)(uint16_t)v.y << 16); o.y = (uint32_t)(uint16_t)v.z | ((uint32_t)(uint16_t)v.w << 16);
      *reinterpret_cast<uint2*>(dst + (size_t)(by + 2 * half + rr) * g.pitch_y + bx) = o;
    }
    return;
  }
  const int cl = s.cls[bi][bj];
  int f[16];
  {
    const int4* cp = reinterpret_cast<const int4*>(&s.coef[cl & 31][cl >> 5][0]);
    const int4 a = cp[0], b = cp[1], c = cp[2], d = cp[3];
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
    f[8] = c.x; f[9] = c.y; f[10] = c.z; f[11] = c.w; f[12] = d.x; f[13] = d.y; f[14] = d.z; f[15] = d.w;
  }
  const bool is7 = sd.alf->luma_filter_7x7 != 0;
#pragma unroll
  for (int rr = 0; rr < 2; rr++) {
    const int trow = 3 + 4 * bi + 2 * half + rr;  // tile row of the output row
    // w[dy+3][i]: samples of tile row trow+dy, tile columns tcx-4 .. tcx+7 (i = 0..11); output px j reads i = j + 4 + dx
    int sum[4];
    if (is7) {
      int r0[12], rp[12], rm[12];
      auto load12 = [&](int* w, int row) {
        const int4* p = reinterpret_cast<const int4*>(&s.t[row][tcx - 4]);
        const int4 a = p[0], b = p[1], c = p[2];
        w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w; w[8] = c.x; w[9] = c.y; w[10] = c.z; w[11] = c.w;
      };
      load12(r0, trow);
#pragma unroll
      for (int j = 0; j < 4; j++)
        sum[j] = f[12] * r0[j + 4] + f[11] * (r0[j + 5] + r0[j + 3]) + f[10] * (r0[j + 6] + r0[j + 2]) + f[9] * (r0[j + 7] + r0[j + 1]);
      load12(rp, trow + 1); load12(rm, trow - 1);
#pragma unroll
      for (int j = 0; j < 4; j++)
        sum[j] += f[4] * (rp[j + 6] + rm[j + 2]) + f[5] * (rp[j + 5] + rm[j + 3]) + f[6] * (rp[j + 4] + rm[j + 4]) +
                  f[7] * (rp[j + 3] + rm[j + 5]) + f[8] * (rp[j + 2] + rm[j + 6]);
      load12(rp, trow + 2); load12(rm, trow - 2);
#pragma unroll
      for (int j = 0; j < 4; j++)
        sum[j] += f[1] * (rp[j + 5] + rm[j + 3]) + f[2] * (rp[j + 4] + rm[j + 4]) + f[3] * (rp[j + 3] + rm[j + 5]);
      {
        const int4 a = *reinterpret_cast<const int4*>(&s.t[trow + 3][tcx]);
        const int4 b = *reinterpret_cast<const int4*>(&s.t[trow - 3][tcx]);
        sum[0] += f[0] * (a.x + b.x); sum[1] += f[0] * (a.y + b.y); sum[2] += f[0] * (a.z + b.z); sum[3] += f[0] * (a.w + b.w);
      }
    } else {
      int r0[12], rp[12], rm[12];
      auto load12 = [&](int* w, int row) {
        const int4* p = reinterpret_cast<const int4*>(&s.t[row][tcx - 4]);
        const int4 a = p[0], b = p[1], c = p[2];
        w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w; w[8] = c.x; w[9] = c.y; w[10] = c.z; w[11] = c.w;
      };
      load12(r0, trow);
#pragma unroll
      for (int j = 0; j < 4; j++) sum[j] = f[6] * r0[j + 4] + f[5] * (r0[j + 5] + r0[j + 3]) + f[4] * (r0[j + 6] + r0[j + 2]);
      load12(rp, trow + 1); load12(rm, trow - 1);
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
    uint2 o; o.x = (uint32_t)o4[0] | ((uint32_t)o4[1] << 16); o.y = (uint32_t)o4[2] | ((uint32_t)o4[3] << 16);
    *reinterpret_cast<uint2*>(dst + (size_t)(by + 2 * half + rr) * g.pitch_y + bx) = o;
  }
}

// ---------------------------------------------------------------------------------------------------------
// Chroma: 5x5 diamond, one filter per picture, no classification.  Tile 64x16 per plane, 4 samples per thread.
// ---------------------------------------------------------------------------------------------------------
constexpr int CT_W = 64, CT_H = 16, CS_W = CT_W + 8, CS_H = CT_H + 4;

__global__ void __launch_bounds__(NT) alf_chroma_kernel(Geom g, const SlotDev* __restrict__ slots, int first_slot, BatchCtl bc) {
  __shared__ __align__(16) int t[CS_H][CS_W];
  const SlotDev& sd = slots[first_slot + (blockIdx.z >> 1)];
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
