// ilf_alf.cu -- adaptive loop filter: 4x4 block classification + 7x7/5x5 luma and 5x5 chroma diamond filters (sm_100a).
//
// Replaces AdaptiveLoopFilter::ALFProcess after coefficient reconstruction (AdaptiveLoopFilter.cpp:89-138):
//   deriveClassificationBlk :292-463   Laplacian activity/direction -> classIdx (0..24), transposeIdx (0..3)
//   filterBlk<7/5>          :465-650   point-symmetric diamond FIR, (sum + 256) >> 9, clip to [0, 2^bd - 1]
// The source is the whole SAO'd picture padded by 3 replicated samples (:90-92, Buffer.h:433-465).  CTUs whose enable
// flag is 0 are copied through (the stage reads one buffer and writes the other).
//
// Data movement (both kernels): band walking over a TMA ring (ilf_ring.cuh).  A CTA owns a band of 32 rows of a plane and
// walks it in tiles of TW = 64 samples; the TMA unit delivers each tile with its halo (box 80 x 38 luma, 80 x 36
// chroma) into a ring of shared-memory stages ahead of the arithmetic.  Per tile the CTA builds a work tile: the staged
// tile plus 4 halo columns from the neighbouring tiles of the ring, with the border padding applied (build_work_tile).
//
// Luma, per tile (2 TW = 128 threads, one thread = one 4x4 block from classification to output):
//   phase 1  work tile (int16, two samples per 32-bit word)
//   phase 2  1-D Laplacians two samples per instruction (|2c - a - b| = max(2c - s, s - 2c) with VIADD.16x2 / VIADDMNMX.S16x2),
//            summed per 2x2 cell; a task walks one word column over 12 rows with a rolling 3-row window
//   phase 3  each thread sums the 4x4 cells of its block's 8x8 window and derives class + transpose (kept in a register)
//   phase 4  each thread filters its block with the (class, transpose) coefficient row of the per-picture table that
//            ilf_set_alf_params precomputed (SlotDev::alf_coef): the 10 window rows are read once from shared memory
//            (3 x 8 bytes each), unpacked once into registers and shared by the block's four output rows
// Tiles whose blocks are all in CTUs with ALF off are copied through without classification.
#include "ilf_common.cuh"
#include "ilf_ring.cuh"

namespace ilf {
namespace {

// ---- band walking (ilf_ring.cuh): shared by the chroma kernel (and the luma kernel) ----
constexpr int TW = ALF_TILE;        // tile width in samples
constexpr int BR = ALF_BAND_ROWS;   // rows of a band
constexpr int WP = TW + 16;         // work-tile pitch in samples: [8: left halo slot][128][8: right halo slot], 288-byte rows
constexpr int WX0 = 8;              // work-tile column of the tile's first sample
#ifndef ALF_RING
#define ALF_RING 4
#endif
#ifndef ALF_L_CTAS
#define ALF_L_CTAS (ALF_TILE_W == 128 ? 3 : 5)
#endif
constexpr int RING_STAGES = ALF_RING;  // the tile being filtered + the tiles in flight
constexpr int L_CTAS = ALF_L_CTAS;     // resident luma CTAs per SM the kernel is built for

// A stage of the ring IS the work tile: the TMA box starts 8 samples left of the tile (16-byte aligned) and is WP wide, so
// it carries the tile's horizontal halo with it; these kernels are instruction-bound, the 12.5 % of re-read columns come
// out of L2 and cost nothing.  Out-of-picture samples arrive as zeros; the filters read the picture "padded by
// replication" (AdaptiveLoopFilter.cpp:90-92, Buffer.h:433-465), so tiles that touch a picture border repeat the nearest
// picture sample into the halo: first along the rows (left / right border), then whole rows (top / bottom border).
// vw = valid samples of this tile's rows, ytop = plane row of staged row 0, ph = plane rows.  CTA-uniform; every thread
// of the CTA must call it.
template <int SR, int HALO, int NTHREADS>
__device__ __forceinline__ void pad_borders(int16_t* __restrict__ W, bool left, bool right, int vw, int ytop, int ph) {
  const bool top = ytop < 0, bottom = ytop + SR > ph;
  if (!(left || right || top || bottom)) return;
  if (left || right) {
    for (int r = threadIdx.x; r < SR; r += NTHREADS) {
      int16_t* wrow = W + r * WP;
      if (left) { const uint32_t e = (uint16_t)wrow[WX0]; *reinterpret_cast<uint2*>(wrow + WX0 - 4) = make_uint2(e | (e << 16), e | (e << 16)); }
      if (right) { const uint32_t e = (uint16_t)wrow[WX0 + vw - 1]; *reinterpret_cast<uint2*>(wrow + WX0 + vw) = make_uint2(e | (e << 16), e | (e << 16)); }
    }
    __syncthreads();
  }
  if (top || bottom) {
    const int r_last = ph - 1 - ytop;  // staged row of the last picture row
    for (int i = threadIdx.x; i < 2 * HALO * (WP / 8); i += NTHREADS) {
      const int q = i / (WP / 8), c = i - q * (WP / 8);
      // q < HALO: rows above the picture <- staged row -ytop; else: the HALO rows below the picture <- staged row r_last
      if (q < HALO) { if (top && q < -ytop) *reinterpret_cast<uint4*>(W + q * WP + 8 * c) = *reinterpret_cast<const uint4*>(W + (-ytop) * WP + 8 * c); }
      else if (bottom && r_last + 1 + (q - HALO) < SR) *reinterpret_cast<uint4*>(W + (r_last + 1 + q - HALO) * WP + 8 * c) = *reinterpret_cast<const uint4*>(W + r_last * WP + 8 * c);
    }
    __syncthreads();
  }
}

constexpr int L_SR = BR + 2 * ALF_HALO_Y;                      // 38 staged rows
constexpr int L_STAGE_BYTES = WP * L_SR * 2;                   // 10944 bytes per box
constexpr int L_STAGE_STRIDE = (L_STAGE_BYTES + 127) & ~127;   // TMA destinations are 128-byte aligned
constexpr int CELL_W = TW / 2 + 2, CELL_H = BR / 2 + 2;        // 66 x 18 cells of 2x2 samples, first cell at (x0-2, y0-2)
constexpr int L_CELL_BYTES = CELL_H * CELL_W * 8;
constexpr int L_SMEM_BYTES = RING_STAGES * L_STAGE_STRIDE + L_CELL_BYTES + RING_STAGES * 8;
constexpr int NT = 2 * TW;                                     // TW / 4 blocks across x 8 block rows
constexpr int BPR = TW / 4;                                    // 4x4 blocks per tile row
constexpr int LAP_ROWS = 6;                                    // sample rows a Laplacian task walks (3 cell rows)
constexpr int LAP_COLS = CELL_W / 2;                           // a task covers two word columns (two cells per cell row)
constexpr int LAP_TASKS = LAP_COLS * (CELL_H * 2 / LAP_ROWS);  // 33 column pairs x 6 row groups
static_assert(CELL_W % 2 == 0 && (CELL_H * 2) % LAP_ROWS == 0 && LAP_TASKS <= 2 * TW, "Laplacian task grid");

__constant__ uint8_t c_th[16] = {0, 1, 2, 2, 2, 2, 2, 3, 3, 3, 3, 3, 3, 3, 3, 4};
__constant__ uint8_t c_transpose[8] = {0, 1, 0, 2, 2, 3, 1, 3};

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

// One work-tile row of a block's window: p = sample x - 4 (8-byte aligned); v[i] = sample x - 4 + i
__device__ __forceinline__ void load_win12(const int16_t* p, int v[12]) {
  const uint2 a = *reinterpret_cast<const uint2*>(p), b = *reinterpret_cast<const uint2*>(p + 4), c = *reinterpret_cast<const uint2*>(p + 8);
  v[0] = a.x & 0xFFFF; v[1] = a.x >> 16; v[2] = a.y & 0xFFFF; v[3] = a.y >> 16;
  v[4] = b.x & 0xFFFF; v[5] = b.x >> 16; v[6] = b.y & 0xFFFF; v[7] = b.y >> 16;
  v[8] = c.x & 0xFFFF; v[9] = c.x >> 16; v[10] = c.y & 0xFFFF; v[11] = c.y >> 16;
}

// Luma CTA: band blockIdx.y, horizontal segment blockIdx.x of nseg.
template <bool CLASSIFY_ONLY>
__device__ __forceinline__ void alf_luma_cta(unsigned char* smem, const Geom& g, const SlotDev& sd, unsigned ctl, int nseg) {
  if (ctl_skip(ctl, 0) || (int)blockIdx.x >= nseg) return;
  const int tid = threadIdx.x;
  const int rows = g.rows;
  const int ntx = (g.width + TW - 1) / TW;
  const int ta = (int)blockIdx.x * ntx / nseg, tb = ((int)blockIdx.x + 1) * ntx / nseg;
  if (ta >= tb) return;
  ring::Walk<RING_STAGES> walk;
  walk.first = ta; walk.last = tb - 1;
  uint2(*cell)[CELL_W] = reinterpret_cast<uint2(*)[CELL_W]>(smem + RING_STAGES * L_STAGE_STRIDE);  // {V | H << 16, D0 | D1 << 16} per 2x2 cell
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + RING_STAGES * L_STAGE_STRIDE + L_CELL_BYTES);
  const int y0 = (int)blockIdx.y * BR;  // local rows
  const int src_buf = ctl_src(ctl, 0);
  int16_t* __restrict__ dst = sd.buf[ctl_dst(ctl, 0)][0];
  const CUtensorMap* map = &sd.tm_alf[0];
  auto stage_ptr = [&](int t) { return reinterpret_cast<int16_t*>(smem + walk.stage(t) * L_STAGE_STRIDE); };
  auto issue = [&](int t) {
    uint64_t* bar = &full[walk.stage(t)];
    ring::mbar_expect_tx(bar, L_STAGE_BYTES);
    ring::tma_load_3d(stage_ptr(t), map, bar, t * TW - WX0, y0 - ALF_HALO_Y, src_buf);
  };
  if (tid == 0) {
    for (int i = 0; i < RING_STAGES; i++) ring::mbar_init(&full[i], 1);
    ring::mbar_init_fence();
  }
  __syncthreads();
  pdl_wait();  // the stage before this one has written the planes read from here on
  if (tid == 0)
    for (int t = walk.first; t <= walk.last && t < walk.first + RING_STAGES; t++) issue(t);

  // this thread's 4x4 block of every tile: BPR blocks across, 8 block rows
  const int bj = tid % BPR, bi = tid / BPR;
  const int by = y0 + 4 * bi;
  const uint8_t* __restrict__ en_row = sd.alf_ctu_enable + (size_t)((by + g.row0) >> g.ctu_log2) * g.ctus_w;
  const int max_val = (1 << g.bd_luma) - 1;
  const bool is7 = CLASSIFY_ONLY ? true : sd.alf->luma_filter_7x7 != 0;
  const int shift = g.bd_luma + 4;
  const uint32_t lap_k = 0x10001u << (g.bd_luma + 1), lap_k2 = lap_k << 1, lap_k4 = lap_k << 2;  // Laplacian lane bias K, 2K, 4K

  for (int tx = ta; tx < tb; tx++) {
    ring::mbar_wait(&full[walk.stage(tx)], walk.parity(tx));
    const int x0 = tx * TW, bx = x0 + 4 * bj;
    const bool blk_in = bx < g.width && by < rows;
    const bool en = blk_in && (CLASSIFY_ONLY || en_row[bx >> g.ctu_log2] != 0);
    int16_t* W = stage_ptr(tx);
    int16_t* out = dst + (size_t)by * g.pitch_y + bx;
    if (!__syncthreads_or(en)) {
      // every CTU under this tile has ALF off: copy through
      if (blk_in) {
#pragma unroll
        for (int o = 0; o < 4; o++) *reinterpret_cast<uint2*>(out + (size_t)o * g.pitch_y) = *reinterpret_cast<const uint2*>(W + (ALF_HALO_Y + 4 * bi + o) * WP + WX0 + 4 * bj);
      }
    } else {
      // ---- phase 1: the stage is the work tile; border tiles get their padding ----
      pad_borders<L_SR, ALF_HALO_Y, NT>(W, tx == 0, tx == ntx - 1, min(TW, g.width - x0), y0 - ALF_HALO_Y, rows);

      // ---- phase 2: Laplacians per 2x2 cell, two samples per instruction.  Task = two word columns (cells 2q, 2q + 1: samples
      //      x0-2+4q .. x0+1+4q) over the 6 sample rows of 3 cell rows: work-tile rows 6 rgp + 1 .. 6 rgp + 6, words 3 + 2q, 4 + 2q.
      //      Everything is plain 32-bit arithmetic on biased lanes (no lane ever borrows or carries): with K = 2^(bd+1),
      //      t' = 2c + K - a - b lies in (0, 2K), and max(t', 2K - t') = K + |2c - a - b|; the 4K a cell collects is taken off
      //      when its two columns are combined.  (__vsub2 / __vneg2 are multi-instruction emulations on sm_100a.) ----
      if (tid < LAP_TASKS) {
        const int q = tid % LAP_COLS, rgp = tid / LAP_COLS;
        const uint32_t* wp = reinterpret_cast<const uint32_t*>(W) + (LAP_ROWS * rgp) * (WP / 2) + 2 + 2 * q;
        // per column j: centre / left-shifted / right-shifted word of the rows above (u), at (m) and below (d)
        uint32_t cu[2], lu[2], ru[2], cm[2], lm[2], rm[2], cd[2], ld[2], rd[2];
        auto load_row = [&](int r, uint32_t (&c)[2], uint32_t (&l)[2], uint32_t (&rr)[2]) {
          const uint2 a = *reinterpret_cast<const uint2*>(wp + r * (WP / 2)), b = *reinterpret_cast<const uint2*>(wp + r * (WP / 2) + 2);
          const uint32_t f12 = __funnelshift_r(a.y, b.x, 16);
          c[0] = a.y; c[1] = b.x; l[0] = __funnelshift_r(a.x, a.y, 16); rr[0] = f12; l[1] = f12; rr[1] = __funnelshift_r(b.x, b.y, 16);
        };
        load_row(0, cu, lu, ru);
        load_row(1, cm, lm, rm);
        uint32_t av[2], ah[2], ad0[2], ad1[2];
#pragma unroll
        for (int i = 0; i < LAP_ROWS; i++) {
          load_row(i + 2, cd, ld, rd);
#pragma unroll
          for (int j = 0; j < 2; j++) {
            const uint32_t c2k = cm[j] + cm[j] + lap_k;
            uint32_t t;
            t = c2k - cu[j] - cd[j]; const uint32_t v = __vmaxu2(t, lap_k2 - t);     // K + |2c - up - down|
            t = c2k - lm[j] - rm[j]; const uint32_t h = __vmaxu2(t, lap_k2 - t);     // K + |2c - left - right|
            t = c2k - lu[j] - rd[j]; const uint32_t d0 = __vmaxu2(t, lap_k2 - t);    // K + |2c - up-left - down-right|
            t = c2k - ru[j] - ld[j]; const uint32_t d1 = __vmaxu2(t, lap_k2 - t);    // K + |2c - up-right - down-left|
            if ((i & 1) == 0) { av[j] = v; ah[j] = h; ad0[j] = d0; ad1[j] = d1; }
            else { av[j] += v; ah[j] += h; ad0[j] += d0; ad1[j] += d1; }
            cu[j] = cm[j]; lu[j] = lm[j]; ru[j] = rm[j]; cm[j] = cd[j]; lm[j] = ld[j]; rm[j] = rd[j];
          }
          if (i & 1) {
            // both columns of each cell: {V, H} and {D0, D1} as 16-bit halves (a cell sum is < 2^15 up to 12 bit)
            uint4 o;
            o.x = __byte_perm(av[0], ah[0], 0x5410) + __byte_perm(av[0], ah[0], 0x7632) - lap_k4;
            o.y = __byte_perm(ad0[0], ad1[0], 0x5410) + __byte_perm(ad0[0], ad1[0], 0x7632) - lap_k4;
            o.z = __byte_perm(av[1], ah[1], 0x5410) + __byte_perm(av[1], ah[1], 0x7632) - lap_k4;
            o.w = __byte_perm(ad0[1], ad1[1], 0x5410) + __byte_perm(ad0[1], ad1[1], 0x7632) - lap_k4;
            *reinterpret_cast<uint4*>(&cell[(LAP_ROWS / 2) * rgp + (i >> 1)][2 * q]) = o;
          }
        }
      }
      __syncthreads();

      // ---- phase 3: 4x4 cells of the block's 8x8 window -> class.  Two cells add without a carry between the 16-bit halves
      //      (12-bit safe); wider sums are taken in 32 bits ----
      int sv = 0, sh = 0, sd0 = 0, sd1 = 0;
#pragma unroll
      for (int r = 0; r < 4; r++) {
        const uint4 c01 = *reinterpret_cast<const uint4*>(&cell[2 * bi + r][2 * bj]);      // cells 0, 1: {vh0, d0, vh1, d1}
        const uint4 c23 = *reinterpret_cast<const uint4*>(&cell[2 * bi + r][2 * bj + 2]);
        const uint32_t a0 = c01.x + c01.z, a1 = c23.x + c23.z, b0 = c01.y + c01.w, b1 = c23.y + c23.w;
        sv += (a0 & 0xFFFF) + (a1 & 0xFFFF); sh += (a0 >> 16) + (a1 >> 16);
        sd0 += (b0 & 0xFFFF) + (b1 & 0xFFFF); sd1 += (b0 >> 16) + (b1 >> 16);
      }
      const int cl = classify(sv, sh, sd0, sd1, shift);
      if (CLASSIFY_ONLY) {
        const int ux = bx >> 2, uy = by >> 2;
        if (blk_in) sd.alf_class[(size_t)uy * g.units_w + ux] = (uint8_t)cl;
      } else if (blk_in) {
        // ---- phase 4: filter the block ----
        const int16_t* wp = W + (4 * bi) * WP + WX0 + 4 * bj - 4;  // window row 0 (= block row 0 minus 3), sample x - 4
        if (!en) {
#pragma unroll
          for (int o = 0; o < 4; o++) *reinterpret_cast<uint2*>(out + (size_t)o * g.pitch_y) = *reinterpret_cast<const uint2*>(wp + (3 + o) * WP + 4);
        } else {
          int f[16];
          {
            const int4* cp = reinterpret_cast<const int4*>(sd.alf_coef + ((cl & 31) * 4 + (cl >> 5)) * 16);
            const int4 a = __ldg(cp), b = __ldg(cp + 1), c = __ldg(cp + 2), d = __ldg(cp + 3);
            f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
            f[8] = c.x; f[9] = c.y; f[10] = c.z; f[11] = c.w; f[12] = d.x; f[13] = d.y; f[14] = d.z; f[15] = d.w;
          }
          // w[s][i] = sample (x - 4 + i, y - 3 + s); output sample j of output row o reads w[o + 3 + dy][j + 4 + dx]
          int w[10][12];
          if (is7) {
#pragma unroll
            for (int s = 0; s < 6; s++) load_win12(wp + s * WP, w[s]);
#pragma unroll
            for (int o = 0; o < 4; o++) {
              load_win12(wp + (o + 6) * WP, w[o + 6]);
              const int(&r0)[12] = w[o + 3];
              int sum[4];
#pragma unroll
              for (int j = 0; j < 4; j++) {
                sum[j] = 256 + f[12] * r0[j + 4] + f[11] * (r0[j + 5] + r0[j + 3]) + f[10] * (r0[j + 6] + r0[j + 2]) + f[9] * (r0[j + 7] + r0[j + 1]);
                sum[j] += f[4] * (w[o + 4][j + 6] + w[o + 2][j + 2]) + f[5] * (w[o + 4][j + 5] + w[o + 2][j + 3]) + f[6] * (w[o + 4][j + 4] + w[o + 2][j + 4]) +
                          f[7] * (w[o + 4][j + 3] + w[o + 2][j + 5]) + f[8] * (w[o + 4][j + 2] + w[o + 2][j + 6]);
                sum[j] += f[1] * (w[o + 5][j + 5] + w[o + 1][j + 3]) + f[2] * (w[o + 5][j + 4] + w[o + 1][j + 4]) + f[3] * (w[o + 5][j + 3] + w[o + 1][j + 5]);
                sum[j] += f[0] * (w[o + 6][j + 4] + w[o][j + 4]);
                sum[j] = __vimin_s32_relu(sum[j] >> 9, max_val);
              }
              const uint32_t p0 = __byte_perm(sum[0], sum[1], 0x5410);
              const uint32_t p1 = __byte_perm(sum[2], sum[3], 0x5410);
              *reinterpret_cast<uint2*>(out + (size_t)o * g.pitch_y) = make_uint2(p0, p1);
            }
          } else {
#pragma unroll
            for (int s = 1; s < 5; s++) load_win12(wp + s * WP, w[s]);
#pragma unroll
            for (int o = 0; o < 4; o++) {
              load_win12(wp + (o + 5) * WP, w[o + 5]);
              const int(&r0)[12] = w[o + 3];
              int sum[4];
#pragma unroll
              for (int j = 0; j < 4; j++) {
                sum[j] = 256 + f[6] * r0[j + 4] + f[5] * (r0[j + 5] + r0[j + 3]) + f[4] * (r0[j + 6] + r0[j + 2]);
                sum[j] += f[1] * (w[o + 4][j + 5] + w[o + 2][j + 3]) + f[2] * (w[o + 4][j + 4] + w[o + 2][j + 4]) + f[3] * (w[o + 4][j + 3] + w[o + 2][j + 5]);
                sum[j] += f[0] * (w[o + 5][j + 4] + w[o + 1][j + 4]);
                sum[j] = __vimin_s32_relu(sum[j] >> 9, max_val);
              }
              const uint32_t p0 = __byte_perm(sum[0], sum[1], 0x5410);
              const uint32_t p1 = __byte_perm(sum[2], sum[3], 0x5410);
              *reinterpret_cast<uint2*>(out + (size_t)o * g.pitch_y) = make_uint2(p0, p1);
            }
          }
        }
      }
    }
    __syncthreads();  // the stage and the cells are free
    if (tid == 0 && tx + RING_STAGES <= walk.last) issue(tx + RING_STAGES);
  }
}

// ---------------------------------------------------------------------------------------------------------
// Chroma: 5x5 diamond, one filter per picture, no classification (filterBlk<ALF_FILTER_5>, AdaptiveLoopFilter.cpp:465-650).
// Band walking over a TMA ring: box (TW + 16) x 36 (2 halo rows each side); a thread filters 8 samples x 2 rows from the work
// tile: 6 window rows of 12 samples are unpacked once into registers and serve both output rows.
// ---------------------------------------------------------------------------------------------------------
constexpr int C_SR = BR + 2 * ALF_HALO_C;
constexpr int C_STAGE_BYTES = WP * C_SR * 2;
static_assert(C_STAGE_BYTES % 128 == 0, "TMA destinations are 128-byte aligned");
constexpr int C_SMEM_BYTES = RING_STAGES * C_STAGE_BYTES + RING_STAGES * 8;
constexpr int NTC = 2 * TW;   // TW / 8 eight-sample groups across x 16 row pairs

// p points at sample x of a work-tile row; v[i] = sample x - 2 + i, i = 0..11
__device__ __forceinline__ void load_row12(const int16_t* p, int v[12]) {
  const uint32_t a = *reinterpret_cast<const uint32_t*>(p - 2);
  const uint4 b = *reinterpret_cast<const uint4*>(p);
  const uint32_t c = *reinterpret_cast<const uint32_t*>(p + 8);
  v[0] = a & 0xFFFF; v[1] = a >> 16;
  v[2] = b.x & 0xFFFF; v[3] = b.x >> 16; v[4] = b.y & 0xFFFF; v[5] = b.y >> 16; v[6] = b.z & 0xFFFF; v[7] = b.z >> 16; v[8] = b.w & 0xFFFF; v[9] = b.w >> 16;
  v[10] = c & 0xFFFF; v[11] = c >> 16;
}

// Chroma CTA: band `cband` (0 .. 2 bands_c - 1: Cb bands, then Cr bands), horizontal segment blockIdx.x of nseg.
__device__ __forceinline__ void alf_chroma_cta(unsigned char* smem, const Geom& g, const SlotDev& sd, unsigned ctl, int cband, int bands_c, int nseg) {
  const int plane = 1 + (cband >= bands_c), band = cband - (plane - 1) * bands_c;
  if (ctl_skip(ctl, plane) || (int)blockIdx.x >= nseg) return;
  const int tid = threadIdx.x;
  const int cw = g.width >> 1, crows = g.rows >> 1;
  const int ntx = (cw + TW - 1) / TW;
  const int ta = (int)blockIdx.x * ntx / nseg, tb = ((int)blockIdx.x + 1) * ntx / nseg;
  if (ta >= tb) return;
  ring::Walk<RING_STAGES> walk;
  walk.first = ta; walk.last = tb - 1;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + RING_STAGES * C_STAGE_BYTES);
  const int by0 = band * BR;
  const int src_buf = ctl_src(ctl, plane);
  int16_t* __restrict__ dst = sd.buf[ctl_dst(ctl, plane)][plane];
  const CUtensorMap* map = &sd.tm_alf[plane];
  auto stage_ptr = [&](int t) { return reinterpret_cast<int16_t*>(smem + walk.stage(t) * C_STAGE_BYTES); };
  auto issue = [&](int t) {
    uint64_t* bar = &full[walk.stage(t)];
    ring::mbar_expect_tx(bar, C_STAGE_BYTES);
    ring::tma_load_3d(stage_ptr(t), map, bar, t * TW - WX0, by0 - ALF_HALO_C, src_buf);
  };
  if (tid == 0) {
    for (int i = 0; i < RING_STAGES; i++) ring::mbar_init(&full[i], 1);
    ring::mbar_init_fence();
  }
  __syncthreads();
  pdl_wait();  // the stage before this one has written the planes read from here on
  if (tid == 0)
    for (int t = walk.first; t <= walk.last && t < walk.first + RING_STAGES; t++) issue(t);

  int f[7];
#pragma unroll
  for (int i = 0; i < 7; i++) f[i] = sd.alf->chroma_coeff[i];
  const int max_val = (1 << g.bd_chroma) - 1;
  const int k = tid % (TW / 8), rg = tid / (TW / 8);  // 8 samples at column 8k, rows 2rg and 2rg + 1 of the band
  const int y = by0 + 2 * rg;
  const uint8_t* __restrict__ en_row = sd.alf_ctu_enable + (size_t)plane * g.ctus_w * g.ctus_h + (size_t)((((y << 1) + g.row0) >> g.ctu_log2) * g.ctus_w);

  for (int tx = ta; tx < tb; tx++) {
    ring::mbar_wait(&full[walk.stage(tx)], walk.parity(tx));
    const int x0 = tx * TW, x = x0 + 8 * k;
    const bool in = x < cw && y < crows;
    const bool en = in && en_row[(x << 1) >> g.ctu_log2] != 0;
    int16_t* W = stage_ptr(tx);
    int16_t* out = dst + (size_t)y * g.pitch_c + x;
    const int nrows = min(2, crows - y);
    if (!__syncthreads_or(en)) {
      // every CTU under this tile has chroma ALF off: copy through
      if (in) {
#pragma unroll
        for (int o = 0; o < 2; o++)
          if (o < nrows) *reinterpret_cast<uint4*>(out + (size_t)o * g.pitch_c) = *reinterpret_cast<const uint4*>(W + (ALF_HALO_C + 2 * rg + o) * WP + WX0 + 8 * k);
      }
    } else {
      pad_borders<C_SR, ALF_HALO_C, NTC>(W, tx == 0, tx == ntx - 1, min(TW, cw - x0), by0 - ALF_HALO_C, crows);
      if (in) {
        const int16_t* wp = W + (2 * rg) * WP + WX0 + 8 * k;  // window row 0 (= output row 0 minus 2), sample x
        if (!en) {
#pragma unroll
          for (int o = 0; o < 2; o++)
            if (o < nrows) *reinterpret_cast<uint4*>(out + (size_t)o * g.pitch_c) = *reinterpret_cast<const uint4*>(wp + (2 + o) * WP);
        } else {
          int w[6][12];
#pragma unroll
          for (int r = 0; r < 6; r++) load_row12(wp + r * WP, w[r]);
#pragma unroll
          for (int o = 0; o < 2; o++) {
            if (o >= nrows) break;
            uint32_t pk2[4];
#pragma unroll
            for (int jj = 0; jj < 4; jj++) {
              int sm[2];
#pragma unroll
              for (int h = 0; h < 2; h++) {
                const int j = 2 * jj + h + 2;  // window index of the output sample
                int sum = 256 + f[6] * w[o + 2][j] + f[5] * (w[o + 2][j + 1] + w[o + 2][j - 1]) + f[4] * (w[o + 2][j + 2] + w[o + 2][j - 2]);
                sum += f[1] * (w[o + 3][j + 1] + w[o + 1][j - 1]) + f[2] * (w[o + 3][j] + w[o + 1][j]) + f[3] * (w[o + 3][j - 1] + w[o + 1][j + 1]);
                sum += f[0] * (w[o + 4][j] + w[o][j]);
                sm[h] = __vimin_s32_relu(sum >> 9, max_val);  // clip to [0, max]
              }
              pk2[jj] = __byte_perm(sm[0], sm[1], 0x5410);
            }
            *reinterpret_cast<uint4*>(out + (size_t)o * g.pitch_c) = make_uint4(pk2[0], pk2[1], pk2[2], pk2[3]);
          }
        }
      }
    }
    __syncthreads();  // the stage is free
    if (tid == 0 && tx + RING_STAGES <= walk.last) issue(tx + RING_STAGES);
  }
}

// One launch for the whole ALF stage: grid rows [0, bands_y) are luma bands, [bands_y, bands_y + 2 bands_c) chroma bands.
// The instruction-bound luma CTAs and the memory-bound chroma CTAs share the SMs instead of running back to back.
// PLANES: 1 = luma only, 2 = chroma only (split launches, ILF_ALF_SPLIT=1), 3 = both.
template <int PLANES>
__global__ void __launch_bounds__(NT, L_CTAS) alf_kernel(Geom g, const SlotDev* __restrict__ slots, int first_slot, BatchCtl bc, int bands_y, int bands_c, int nseg_y, int nseg_c) {
  extern __shared__ __align__(128) unsigned char smem[];
  pdl_launch_dependents();
  const SlotDev& sd = slots[first_slot + bc.slot[blockIdx.z]];
  const unsigned ctl = bc.v[blockIdx.z];
  if (PLANES == 1 || (PLANES == 3 && (int)blockIdx.y < bands_y)) alf_luma_cta<false>(smem, g, sd, ctl, nseg_y);
  else alf_chroma_cta(smem, g, sd, ctl, (int)blockIdx.y - (PLANES == 3 ? bands_y : 0), bands_c, nseg_c);
}
__global__ void __launch_bounds__(NT, L_CTAS) alf_classify_kernel(Geom g, const SlotDev* __restrict__ slots, int first_slot, BatchCtl bc, int nseg) {
  extern __shared__ __align__(128) unsigned char smem[];
  pdl_launch_dependents();
  alf_luma_cta<true>(smem, g, slots[first_slot + bc.slot[blockIdx.z]], bc.v[blockIdx.z], nseg);
}
static_assert(NT == NTC, "luma and chroma CTAs share one launch");

}  // namespace

static int alf_nseg(int bands_total, int ntx, bool luma) {
  if (luma) return pick_segments(bands_total, ntx, 148 * L_CTAS);
  int nseg = (148 * L_CTAS + bands_total - 1) / bands_total;
  return nseg < 1 ? 1 : (nseg > ntx ? ntx : nseg);
}

// planes: bit 0 = luma, bit 1 = chroma (the slots' control words say which planes of which slot really run)
void launch_alf(const Geom& g, const SlotDev* slots, int first_slot, int num_slots, const BatchCtl& ctl, int planes, cudaStream_t st) {
  static bool attr_set[64] = {};
  if (first_launch_on_device(attr_set)) {
    cudaFuncSetAttribute(alf_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, L_SMEM_BYTES);
    cudaFuncSetAttribute(alf_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, C_SMEM_BYTES);
    cudaFuncSetAttribute(alf_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, L_SMEM_BYTES > C_SMEM_BYTES ? L_SMEM_BYTES : C_SMEM_BYTES);
  }
  const int bands_y = (g.rows + BR - 1) / BR, bands_c = (g.rows / 2 + BR - 1) / BR;
  const int nseg_y = alf_nseg(bands_y * num_slots, (g.width + TW - 1) / TW, true);
  const int nseg_c = alf_nseg(2 * bands_c * num_slots, (g.width / 2 + TW - 1) / TW, false);
  if (planes == 1) launch_pdl(alf_kernel<1>, dim3(nseg_y, bands_y, num_slots), dim3(NT), L_SMEM_BYTES, st, g, slots, first_slot, ctl, bands_y, bands_c, nseg_y, nseg_c);
  else if (planes == 2) launch_pdl(alf_kernel<2>, dim3(nseg_c, 2 * bands_c, num_slots), dim3(NT), C_SMEM_BYTES, st, g, slots, first_slot, ctl, bands_y, bands_c, nseg_y, nseg_c);
  else launch_pdl(alf_kernel<3>, dim3(nseg_y > nseg_c ? nseg_y : nseg_c, bands_y + 2 * bands_c, num_slots), dim3(NT), L_SMEM_BYTES > C_SMEM_BYTES ? L_SMEM_BYTES : C_SMEM_BYTES, st, g,
                  slots, first_slot, ctl, bands_y, bands_c, nseg_y, nseg_c);
}

void launch_alf_classify(const Geom& g, const SlotDev* slots, int first_slot, int num_slots, const BatchCtl& ctl, cudaStream_t st) {
  static bool attr_set[64] = {};
  if (first_launch_on_device(attr_set)) { cudaFuncSetAttribute(alf_classify_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, L_SMEM_BYTES); }
  const int bands_y = (g.rows + BR - 1) / BR;
  const int nseg = alf_nseg(bands_y * num_slots, (g.width + TW - 1) / TW, true);
  launch_pdl(alf_classify_kernel, dim3(nseg, bands_y, num_slots), dim3(NT), L_SMEM_BYTES, st, g, slots, first_slot, ctl, nseg);
}

}  // namespace ilf
