/*
 * oracle/ilf_oracle.c -- TEST INFRASTRUCTURE: CPU restatement of VTM 2.1's in-loop filter chain.
 *
 * Plain C, single-threaded, written for clarity.  It consumes ONLY the flat side information of
 * include/ilf_b200.h (the same arrays the CUDA library gets), never VTM objects, and is the checker
 * the CUDA path is compared against.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline
 * leg may load this library; the product (vvcsoftware_vtm_b200/) never does.
 *
 * PARITY PINNING: the reference ships no golden vectors.  This restatement is
 * pinned against outputs of the reference itself: oracle/_ref/vtm_capture (the unmodified reference
 * decoder with a dump hook, oracle/capture_hook.cpp) writes the picture before deblocking and after each
 * stage for real bitstreams; tests/test_oracle_vs_reference.py requires bit-equality on every captured
 * picture (committed fixtures under tests/golden/ plus fresh captures when oracle/_ref is present), and
 * tests/test_oracle_units.py drives the reference's own SAO/ALF block functions (oracle/_ref/libvtm_units.so)
 * with random blocks and availability flags.
 *
 * Each function cites the reference lines it follows (paths relative to /root/reference/source/Lib/CommonLib).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "ilf_b200.h"

typedef int16_t Pel;

static inline int clip3(int lo, int hi, int v) { return v < lo ? lo : (v > hi ? hi : v); }
static inline int iabs(int v) { return v < 0 ? -v : v; }

/* ------------------------------------------------------------------------------------------------
 * Deblocking
 * ---------------------------------------------------------------------------------------------- */

/* LoopFilter.cpp:66-80 (MAX_QP = 63, CommonDef.h:140) */
static const uint8_t kTc[66] = {0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  1,  1,  1,  1,
                                1,  1,  1,  1,  1,  2,  2,  2,  2,  3,  3,  3,  3,  4,  4,  4,  5,  5,  6,  6,  7,  8,
                                9,  10, 11, 13, 14, 16, 18, 20, 22, 24, 26, 28, 30, 32, 34, 36, 38, 40, 42, 44, 46, 48};
static const uint8_t kBeta[64] = {0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  6,  7,  8,  9,  10, 11,
                                  12, 13, 14, 15, 16, 17, 18, 20, 22, 24, 26, 28, 30, 32, 34, 36, 38, 40, 42, 44, 46, 48,
                                  50, 52, 54, 56, 58, 60, 62, 64, 66, 68, 70, 72, 74, 76, 78, 80, 82, 84, 86, 88};
/* Rom.cpp:523-530, row CHROMA_420; chromaQPMappingTableSize = MAX_QP + 7 = 70 (Rom.h:98) */
static const uint8_t kChromaScale420[70] = {0,  1,  2,  3,  4,  5,  6,  7,  8,  9,  10, 11, 12, 13, 14, 15, 16, 17,
                                            18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 29, 30, 31, 32, 33, 33,
                                            34, 34, 35, 35, 36, 36, 37, 37, 38, 39, 40, 41, 42, 43, 44, 45, 46, 47,
                                            48, 49, 50, 51, 52, 53, 54, 55, 56, 57, 58, 59, 60, 61, 62, 63};

typedef struct {
  int w, h, units_w, units_h, ctu_log2, ctus_w;
  int bd_luma, bd_chroma;
  const ilf_deblock_params* p;
  const uint32_t* info;   /* luma layer */
  const uint32_t* info_c; /* layer used for chroma edges (== info when there is no chroma tree) */
  const int16_t* mv16;
  const int32_t* mv32;
  const uint8_t* ctu_slice;
} db_ctx;

static inline int mv_comp(const db_ctx* c, int unit, int k) {
  if (c->mv32) return c->mv32[(size_t)unit * 4 + k];
  if (c->mv16) return c->mv16[(size_t)unit * 4 + k];
  return 0;
}

/* xGetBoundaryStrengthSingle, LoopFilter.cpp:419-541.  q/p = unit indices on the Q and P side. */
static int boundary_strength(const db_ctx* c, const uint32_t* info, int q, int p, uint32_t tu_bit) {
  const uint32_t iq = info[q], ip = info[p];
  if ((iq | ip) & ILF_BI_INTRA) return 2; /* :433 */
  if ((iq & tu_bit) && ((iq | ip) & ILF_BI_CBF)) return 1; /* :444 */
  const int thr = c->p->mv_threshold;
  const int rq0 = (iq >> 16) & 0xFF, rq1 = (iq >> 24) & 0xFF;
  const int rp0 = (ip >> 16) & 0xFF, rp1 = (ip >> 24) & 0xFF;
  const int q0x = mv_comp(c, q, 0), q0y = mv_comp(c, q, 1), q1x = mv_comp(c, q, 2), q1y = mv_comp(c, q, 3);
  const int p0x = mv_comp(c, p, 0), p0y = mv_comp(c, p, 1), p1x = mv_comp(c, p, 2), p1y = mv_comp(c, p, 3);
  if ((iq | ip) & ILF_BI_BSLICE) { /* :454-514 */
    if ((rp0 == rq0 && rp1 == rq1) || (rp0 == rq1 && rp1 == rq0)) {
      const int d00 = iabs(q0x - p0x) >= thr || iabs(q0y - p0y) >= thr; /* Q0 vs P0 */
      const int d11 = iabs(q1x - p1x) >= thr || iabs(q1y - p1y) >= thr; /* Q1 vs P1 */
      const int d10 = iabs(q1x - p0x) >= thr || iabs(q1y - p0y) >= thr; /* Q1 vs P0 */
      const int d01 = iabs(q0x - p1x) >= thr || iabs(q0y - p1y) >= thr; /* Q0 vs P1 */
      if (rp0 != rp1) {
        if (rp0 == rq0) return (d00 || d11) ? 1 : 0;
        return (d10 || d01) ? 1 : 0;
      }
      return ((d00 || d11) && (d10 || d01)) ? 1 : 0;
    }
    return 1;
  }
  /* P slice, :517-540: list 0 only */
  if (rp0 != rq0) return 1;
  return (iabs(q0x - p0x) >= thr || iabs(q0y - p0y) >= thr) ? 1 : 0;
}

/* xPelFilterLuma, LoopFilter.cpp:856-916.  `s` points at q0, `o` is the step across the edge. */
static void pel_filter_luma(Pel* s, ptrdiff_t o, int tc, int sw, int no_p, int no_q, int thr_cut, int second_p,
                            int second_q, int max_val) {
  const int m4 = s[0], m3 = s[-o], m5 = s[o], m2 = s[-2 * o], m6 = s[2 * o], m1 = s[-3 * o], m7 = s[3 * o],
            m0 = s[-4 * o];
  if (sw) {
    s[-o] = (Pel)clip3(m3 - 2 * tc, m3 + 2 * tc, (m1 + 2 * m2 + 2 * m3 + 2 * m4 + m5 + 4) >> 3);
    s[0] = (Pel)clip3(m4 - 2 * tc, m4 + 2 * tc, (m2 + 2 * m3 + 2 * m4 + 2 * m5 + m6 + 4) >> 3);
    s[-2 * o] = (Pel)clip3(m2 - 2 * tc, m2 + 2 * tc, (m1 + m2 + m3 + m4 + 2) >> 2);
    s[o] = (Pel)clip3(m5 - 2 * tc, m5 + 2 * tc, (m3 + m4 + m5 + m6 + 2) >> 2);
    s[-3 * o] = (Pel)clip3(m1 - 2 * tc, m1 + 2 * tc, (2 * m0 + 3 * m1 + m2 + m3 + m4 + 4) >> 3);
    s[2 * o] = (Pel)clip3(m6 - 2 * tc, m6 + 2 * tc, (m3 + m4 + m5 + 3 * m6 + 2 * m7 + 4) >> 3);
  } else {
    int delta = (9 * (m4 - m3) - 3 * (m5 - m2) + 8) >> 4;
    if (iabs(delta) < thr_cut) {
      delta = clip3(-tc, tc, delta);
      s[-o] = (Pel)clip3(0, max_val, m3 + delta);
      s[0] = (Pel)clip3(0, max_val, m4 - delta);
      const int tc2 = tc >> 1;
      if (second_p) {
        const int d1 = clip3(-tc2, tc2, (((m1 + m3 + 1) >> 1) - m2 + delta) >> 1);
        s[-2 * o] = (Pel)clip3(0, max_val, m2 + d1);
      }
      if (second_q) {
        const int d2 = clip3(-tc2, tc2, (((m6 + m4 + 1) >> 1) - m5 - delta) >> 1);
        s[o] = (Pel)clip3(0, max_val, m5 + d2);
      }
    }
  }
  if (no_p) { s[-o] = (Pel)m3; s[-2 * o] = (Pel)m2; s[-3 * o] = (Pel)m1; }
  if (no_q) { s[0] = (Pel)m4; s[o] = (Pel)m5; s[2 * o] = (Pel)m6; }
}

static inline int calc_dp(const Pel* s, ptrdiff_t o) { return iabs(s[-3 * o] - 2 * s[-2 * o] + s[-o]); }  /* :972 */
static inline int calc_dq(const Pel* s, ptrdiff_t o) { return iabs(s[0] - 2 * s[o] + s[2 * o]); }          /* :977 */
static inline int use_strong(const Pel* s, ptrdiff_t o, int d, int beta, int tc) {                         /* :960 */
  const int d_strong = iabs(s[-4 * o] - s[-o]) + iabs(s[3 * o] - s[0]);
  return (d_strong < (beta >> 3)) && (d < (beta >> 2)) && (iabs(s[-o] - s[0]) < ((tc * 5 + 1) >> 1));
}

static inline const ilf_slice_params* slice_of(const db_ctx* c, int ux, int uy) {
  int s = 0;
  if (c->ctu_slice) s = c->ctu_slice[(size_t)((uy * 4) >> c->ctu_log2) * c->ctus_w + ((ux * 4) >> c->ctu_log2)];
  return &c->p->slices[s];
}

/* One 4-sample segment of a luma edge: the body of the iIdx loop of xEdgeFilterLuma, LoopFilter.cpp:596-680. */
static void edge_luma_unit(const db_ctx* c, Pel* y, ptrdiff_t stride, int dir, int ux, int uy) {
  const int q = uy * c->units_w + ux;
  const int p = dir == 0 ? q - 1 : q - c->units_w;
  const uint32_t iq = c->info[q], ip = c->info[p];
  if (!(iq & (dir == 0 ? ILF_BI_EDGE_V : ILF_BI_EDGE_H))) return;
  const int bs = boundary_strength(c, c->info, q, p, dir == 0 ? ILF_BI_TU_V : ILF_BI_TU_H);
  if (!bs) return;
  const ilf_slice_params* sl = slice_of(c, ux, uy); /* cu.slice = slice of the Q side (:554) */
  const int qp = ((int8_t)(ip >> 8) + (int8_t)(iq >> 8) + 1) >> 1; /* :626 */
  const int idx_tc = clip3(0, 63 + 2, qp + 2 * (bs - 1) + (sl->tc_offset_div2 << 1));
  const int idx_b = clip3(0, 63, qp + (sl->beta_offset_div2 << 1));
  const int scale = 1 << (c->bd_luma - 8);
  const int tc = kTc[idx_tc] * scale, beta = kBeta[idx_b] * scale;
  const int side_thr = (beta + (beta >> 1)) >> 3, thr_cut = tc * 10;
  const int no_p = (ip & ILF_BI_NOFILT) != 0, no_q = (iq & ILF_BI_NOFILT) != 0;
  const ptrdiff_t o = dir == 0 ? 1 : stride;      /* across the edge */
  const ptrdiff_t step = dir == 0 ? stride : 1;   /* along the edge  */
  Pel* s = y + (ptrdiff_t)(uy * 4) * stride + ux * 4;
  const int dp0 = calc_dp(s, o), dq0 = calc_dq(s, o), dp3 = calc_dp(s + 3 * step, o), dq3 = calc_dq(s + 3 * step, o);
  const int d0 = dp0 + dq0, d3 = dp3 + dq3, dp = dp0 + dp3, dq = dq0 + dq3, d = d0 + d3;
  if (d < beta) {
    const int fp = dp < side_thr, fq = dq < side_thr;
    const int sw = use_strong(s, o, 2 * d0, beta, tc) && use_strong(s + 3 * step, o, 2 * d3, beta, tc);
    for (int i = 0; i < 4; i++) pel_filter_luma(s + i * step, o, tc, sw, no_p, no_q, thr_cut, fp, fq, (1 << c->bd_luma) - 1);
  }
}

/* One 4-luma-sample (2 chroma lines) segment of a chroma edge: xEdgeFilterChroma, LoopFilter.cpp:684-838. */
static void edge_chroma_unit(const db_ctx* c, Pel* cb, Pel* cr, ptrdiff_t stride, int dir, int ux, int uy) {
  if (((dir == 0 ? ux : uy) & 3) != 0) return; /* only edges on the 16-luma-sample grid (:713-724) */
  const int q = uy * c->units_w + ux;
  const int p = dir == 0 ? q - 1 : q - c->units_w;
  const uint32_t iq = c->info_c[q], ip = c->info_c[p];
  if (!(iq & (dir == 0 ? ILF_BI_EDGE_V : ILF_BI_EDGE_H))) return;
  const int bs = boundary_strength(c, c->info_c, q, p, dir == 0 ? ILF_BI_TU_V : ILF_BI_TU_H);
  if (bs < 2) return; /* :769 */
  const ilf_slice_params* sl = slice_of(c, ux, uy);
  const int no_p = (ip & ILF_BI_NOFILT) != 0, no_q = (iq & ILF_BI_NOFILT) != 0;
  const int max_val = (1 << c->bd_chroma) - 1;
  const ptrdiff_t o = dir == 0 ? 1 : stride, step = dir == 0 ? stride : 1;
  for (int comp = 0; comp < 2; comp++) {
    Pel* s = (comp == 0 ? cb : cr) + (ptrdiff_t)(uy * 2) * stride + ux * 2;
    int qp = (((int8_t)(ip >> 8) + (int8_t)(iq >> 8) + 1) >> 1) + (comp == 0 ? c->p->cb_qp_offset : c->p->cr_qp_offset);
    if (qp >= 70) qp -= 6;                    /* :812-817 (4:2:0) */
    else if (qp >= 0) qp = kChromaScale420[qp]; /* :823-826 */
    const int idx_tc = clip3(0, 63 + 2, qp + 2 * (bs - 1) + (sl->tc_offset_div2 << 1));
    const int tc = kTc[idx_tc] * (1 << (c->bd_chroma - 8));
    for (int i = 0; i < 2; i++) { /* xPelFilterChroma, :928-949 */
      Pel* t = s + i * step;
      const int m4 = t[0], m3 = t[-o], m5 = t[o], m2 = t[-2 * o];
      const int delta = clip3(-tc, tc, (((m4 - m3) << 2) + m2 - m5 + 4) >> 3);
      if (!no_p) t[-o] = (Pel)clip3(0, max_val, m3 + delta);
      if (!no_q) t[0] = (Pel)clip3(0, max_val, m4 - delta);
    }
  }
}

/* LoopFilter::loopFilterPic, LoopFilter.cpp:149-230: all vertical edges of the picture, then all horizontal
 * edges on the result.  In place on y/cb/cr (strides in samples). */
int ilf_oracle_deblock(int16_t* y, ptrdiff_t sy, int16_t* cb, int16_t* cr, ptrdiff_t sc, int width, int height,
                       int bd_luma, int bd_chroma, int ctu_log2, const ilf_deblock_params* params,
                       const uint32_t* info, const uint32_t* info_chroma, const int16_t* mv16, const int32_t* mv32,
                       const uint8_t* ctu_slice) {
  db_ctx c;
  c.w = width; c.h = height; c.units_w = width / 4; c.units_h = height / 4; c.ctu_log2 = ctu_log2;
  c.ctus_w = (width + (1 << ctu_log2) - 1) >> ctu_log2;
  c.bd_luma = bd_luma; c.bd_chroma = bd_chroma; c.p = params; c.info = info;
  c.info_c = info_chroma ? info_chroma : info; c.mv16 = mv16; c.mv32 = mv32; c.ctu_slice = ctu_slice;
  for (int dir = 0; dir < 2; dir++)
    for (int uy = 0; uy < c.units_h; uy++)
      for (int ux = 0; ux < c.units_w; ux++) {
        if ((dir == 0 ? ux : uy) == 0) continue; /* picture border is never an edge (:414-415) */
        edge_luma_unit(&c, y, sy, dir, ux, uy);
        edge_chroma_unit(&c, cb, cr, sc, dir, ux, uy);
      }
  return 0;
}

/* bS map for diagnostics/tests: out[dir][uy][ux] = bS (0..2) where the edge flag is set, else 0. */
int ilf_oracle_bs_map(uint8_t* out, int width, int height, const ilf_deblock_params* params, const uint32_t* info,
                      const int16_t* mv16, const int32_t* mv32) {
  db_ctx c;
  memset(&c, 0, sizeof(c));
  c.units_w = width / 4; c.units_h = height / 4; c.p = params; c.info = info; c.info_c = info; c.mv16 = mv16; c.mv32 = mv32;
  const size_t n = (size_t)c.units_w * c.units_h;
  memset(out, 0, 2 * n);
  for (int dir = 0; dir < 2; dir++)
    for (int uy = 0; uy < c.units_h; uy++)
      for (int ux = 0; ux < c.units_w; ux++) {
        if ((dir == 0 ? ux : uy) == 0) continue;
        const int q = uy * c.units_w + ux, p = dir == 0 ? q - 1 : q - c.units_w;
        if (!(info[q] & (dir == 0 ? ILF_BI_EDGE_V : ILF_BI_EDGE_H))) continue;
        out[dir * n + q] = (uint8_t)boundary_strength(&c, info, q, p, dir == 0 ? ILF_BI_TU_V : ILF_BI_TU_H);
      }
  return 0;
}

/* ------------------------------------------------------------------------------------------------
 * SAO
 * ---------------------------------------------------------------------------------------------- */
static inline int sgn(int v) { return (v > 0) - (v < 0); } /* SampleAdaptiveOffset.h:58-61 */

/* offsetBlock, SampleAdaptiveOffset.cpp:292-508, restated per sample.  The row/column special cases of the
 * reference (:308-487) say: an edge-offset sample is modified iff BOTH its neighbours along the class direction
 * lie inside the CTU block or inside a neighbouring CTU whose availability flag is set; otherwise it keeps the
 * source value.  (x, y) are block-relative; bw x bh is the (possibly clipped) block of this component. */
static int nb_available(int x, int y, int bw, int bh, unsigned avail) {
  const int l = x < 0, r = x >= bw, a = y < 0, b = y >= bh;
  if (!l && !r && !a && !b) return 1;
  if (a && l) return (avail & ILF_AVAIL_AL) != 0;
  if (a && r) return (avail & ILF_AVAIL_AR) != 0;
  if (b && l) return (avail & ILF_AVAIL_BL) != 0;
  if (b && r) return (avail & ILF_AVAIL_BR) != 0;
  if (l) return (avail & ILF_AVAIL_L) != 0;
  if (r) return (avail & ILF_AVAIL_R) != 0;
  if (a) return (avail & ILF_AVAIL_A) != 0;
  return (avail & ILF_AVAIL_B) != 0;
}

static void sao_block(const Pel* src, ptrdiff_t ss, Pel* dst, ptrdiff_t ds, int bw, int bh, int type, int band_pos,
                      const int16_t off4[4], unsigned avail, int bit_depth) {
  const int max_val = (1 << bit_depth) - 1;
  if (type == ILF_SAO_BO) { /* :489-501 */
    const int shift = bit_depth - 5;
    for (int y = 0; y < bh; y++)
      for (int x = 0; x < bw; x++) {
        const int v = src[y * ss + x];
        const int k = ((v >> shift) - band_pos) & 31;
        dst[y * ds + x] = (Pel)clip3(0, max_val, v + (k < 4 ? off4[k] : 0));
      }
    return;
  }
  static const int dx[4] = {1, 0, 1, -1}, dy[4] = {0, 1, 1, 1}; /* EO_0, EO_90, EO_135, EO_45: second neighbour b = c + (dx,dy), first a = c - (dx,dy) */
  const int ox = dx[type], oy = dy[type];
  const int off5[5] = {off4[0], off4[1], 0, off4[2], off4[3]};
  for (int y = 0; y < bh; y++)
    for (int x = 0; x < bw; x++) {
      if (!nb_available(x - ox, y - oy, bw, bh, avail) || !nb_available(x + ox, y + oy, bw, bh, avail)) continue;
      const int c = src[y * ss + x];
      const int e = sgn(c - src[(y - oy) * ss + (x - ox)]) + sgn(c - src[(y + oy) * ss + (x + ox)]);
      dst[y * ds + x] = (Pel)clip3(0, max_val, c + off5[2 + e]);
    }
}

/* SAOProcess after parameter resolution, SampleAdaptiveOffset.cpp:585-601 + offsetCTU :510-562.
 * src = deblocked picture (the reference copies it to m_tempBuf, :587), dst = same picture modified in place
 * by the reference; here dst must hold a copy of src on entry (samples of CTUs with SAO off stay untouched). */
int ilf_oracle_sao(const int16_t* const src[3], const ptrdiff_t sstride[3], int16_t* const dst[3],
                   const ptrdiff_t dstride[3], int width, int height, int bd_luma, int bd_chroma, int ctu_log2,
                   const ilf_sao_ctu* ctus) {
  const int ctu = 1 << ctu_log2, cw = (width + ctu - 1) >> ctu_log2, ch = (height + ctu - 1) >> ctu_log2;
  for (int cy = 0; cy < ch; cy++)
    for (int cx = 0; cx < cw; cx++) {
      const ilf_sao_ctu* p = &ctus[cy * cw + cx];
      const int x0 = cx << ctu_log2, y0 = cy << ctu_log2;
      const int bw = (x0 + ctu > width) ? width - x0 : ctu, bh = (y0 + ctu > height) ? height - y0 : ctu;
      for (int c = 0; c < 3; c++) {
        if (p->type[c] == ILF_SAO_OFF) continue;
        const int sh = c ? 1 : 0;
        sao_block(src[c] + (ptrdiff_t)(y0 >> sh) * sstride[c] + (x0 >> sh), sstride[c],
                  dst[c] + (ptrdiff_t)(y0 >> sh) * dstride[c] + (x0 >> sh), dstride[c], bw >> sh, bh >> sh, p->type[c],
                  p->band_pos[c], p->offset[c], p->avail, c ? bd_chroma : bd_luma);
      }
    }
  return 0;
}

/* ------------------------------------------------------------------------------------------------
 * ALF
 * ---------------------------------------------------------------------------------------------- */
/* Sample fetch with the 3-sample replicate padding of the picture (extendBorderPel, Buffer.h:433-465 via
 * AdaptiveLoopFilter.cpp:90-92) expressed as coordinate clamping. */
static inline int px(const Pel* p, ptrdiff_t s, int w, int h, int x, int y) {
  x = x < 0 ? 0 : (x >= w ? w - 1 : x);
  y = y < 0 ? 0 : (y >= h ? h - 1 : y);
  return p[(ptrdiff_t)y * s + x];
}

/* deriveClassificationBlk, AdaptiveLoopFilter.cpp:292-463, for ONE 4x4 block at (bx, by): Laplacians of every
 * sample of the 8x8 window [bx-2, bx+6) x [by-2, by+6) (the reference accumulates them per 2x2 cell :338-341 and
 * then over 4x4 cells :343-353, :385-388, which is the same sum).  Returns classIdx | transposeIdx << 5. */
static uint8_t alf_classify_block(const Pel* p, ptrdiff_t s, int w, int h, int bx, int by, int shift) {
  static const int th[16] = {0, 1, 2, 2, 2, 2, 2, 3, 3, 3, 3, 3, 3, 3, 3, 4};
  static const int transpose_table[8] = {0, 1, 0, 2, 2, 3, 1, 3};
  int sum_v = 0, sum_h = 0, sum_d0 = 0, sum_d1 = 0;
  for (int y = by - 2; y < by + 6; y++)
    for (int x = bx - 2; x < bx + 6; x++) {
      const int c2 = px(p, s, w, h, x, y) << 1;
      sum_v += iabs(c2 - px(p, s, w, h, x, y - 1) - px(p, s, w, h, x, y + 1));
      sum_h += iabs(c2 - px(p, s, w, h, x - 1, y) - px(p, s, w, h, x + 1, y));
      sum_d0 += iabs(c2 - px(p, s, w, h, x - 1, y - 1) - px(p, s, w, h, x + 1, y + 1));
      sum_d1 += iabs(c2 - px(p, s, w, h, x + 1, y - 1) - px(p, s, w, h, x - 1, y + 1));
    }
  const int activity = clip3(0, 15, ((sum_v + sum_h) * 32) >> shift);
  int class_idx = th[activity];
  int hv1, hv0, d1, d0, dir_hv, dir_d, hvd1, hvd0, main_dir, sec_dir;
  if (sum_v > sum_h) { hv1 = sum_v; hv0 = sum_h; dir_hv = 1; } else { hv1 = sum_h; hv0 = sum_v; dir_hv = 3; }
  if (sum_d0 > sum_d1) { d1 = sum_d0; d0 = sum_d1; dir_d = 0; } else { d1 = sum_d1; d0 = sum_d0; dir_d = 2; }
  /* :420 multiplies in `int`; x86 imul / _mm_mullo_epi32 wrap mod 2^32, so do we. */
  if ((int32_t)((uint32_t)d1 * (uint32_t)hv0) > (int32_t)((uint32_t)hv1 * (uint32_t)d0)) {
    hvd1 = d1; hvd0 = d0; main_dir = dir_d; sec_dir = dir_hv;
  } else {
    hvd1 = hv1; hvd0 = hv0; main_dir = dir_hv; sec_dir = dir_d;
  }
  int strength = 0;
  if (hvd1 > 2 * hvd0) strength = 1;
  if (hvd1 * 2 > 9 * hvd0) strength = 2;
  if (strength) class_idx += (((main_dir & 1) << 1) + strength) * 5;
  return (uint8_t)(class_idx | (transpose_table[main_dir * 2 + (sec_dir >> 1)] << 5));
}

/* deriveClassification over a whole picture (AdaptiveLoopFilter.cpp:274-290); out[units_h][units_w]. */
int ilf_oracle_alf_classify(const int16_t* y, ptrdiff_t stride, int width, int height, int bd_luma, uint8_t* out) {
  for (int by = 0; by < height; by += 4)
    for (int bx = 0; bx < width; bx += 4)
      out[(size_t)(by / 4) * (width / 4) + bx / 4] = alf_classify_block(y, stride, width, height, bx, by, bd_luma + 4);
  return 0;
}

/* Coefficient order after transposition, AdaptiveLoopFilter.cpp:541-575. */
static void alf_transpose_coeff(const int16_t* c, int t, int is7, int f[13]) {
  static const uint8_t p7[4][13] = {{0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12},
                                    {9, 4, 10, 8, 1, 5, 11, 7, 3, 0, 2, 6, 12},
                                    {0, 3, 2, 1, 8, 7, 6, 5, 4, 9, 10, 11, 12},
                                    {9, 8, 10, 4, 3, 7, 11, 5, 1, 0, 2, 6, 12}};
  static const uint8_t p5[4][7] = {{0, 1, 2, 3, 4, 5, 6}, {4, 1, 5, 3, 0, 2, 6}, {0, 3, 2, 1, 4, 5, 6}, {4, 3, 5, 1, 0, 2, 6}};
  if (is7) for (int i = 0; i < 13; i++) f[i] = c[p7[t][i]];
  else     for (int i = 0; i < 7; i++) f[i] = c[p5[t][i]];
}

/* One sample of filterBlk<ALF_FILTER_7/5>, AdaptiveLoopFilter.cpp:590-626. */
static inline int alf_sample(const Pel* p, ptrdiff_t s, int w, int h, int x, int y, const int f[13], int is7, int max_val) {
#define P(dx, dy) px(p, s, w, h, x + (dx), y + (dy))
  int sum = 0;
  if (is7) {
    sum += f[0] * (P(0, 3) + P(0, -3));
    sum += f[1] * (P(1, 2) + P(-1, -2));
    sum += f[2] * (P(0, 2) + P(0, -2));
    sum += f[3] * (P(-1, 2) + P(1, -2));
    sum += f[4] * (P(2, 1) + P(-2, -1));
    sum += f[5] * (P(1, 1) + P(-1, -1));
    sum += f[6] * (P(0, 1) + P(0, -1));
    sum += f[7] * (P(-1, 1) + P(1, -1));
    sum += f[8] * (P(-2, 1) + P(2, -1));
    sum += f[9] * (P(3, 0) + P(-3, 0));
    sum += f[10] * (P(2, 0) + P(-2, 0));
    sum += f[11] * (P(1, 0) + P(-1, 0));
    sum += f[12] * P(0, 0);
  } else {
    sum += f[0] * (P(0, 2) + P(0, -2));
    sum += f[1] * (P(1, 1) + P(-1, -1));
    sum += f[2] * (P(0, 1) + P(0, -1));
    sum += f[3] * (P(-1, 1) + P(1, -1));
    sum += f[4] * (P(2, 0) + P(-2, 0));
    sum += f[5] * (P(1, 0) + P(-1, 0));
    sum += f[6] * P(0, 0);
  }
#undef P
  return clip3(0, max_val, (sum + 256) >> 9);
}

/* ALFProcess after coefficient reconstruction, AdaptiveLoopFilter.cpp:89-138.  src = SAO output, dst holds a
 * copy of src on entry (CTUs with the flag off stay untouched). */
int ilf_oracle_alf(const int16_t* const src[3], const ptrdiff_t sstride[3], int16_t* const dst[3],
                   const ptrdiff_t dstride[3], int width, int height, int bd_luma, int bd_chroma, int ctu_log2,
                   const ilf_alf_params* params, const uint8_t* ctu_enable) {
  const int ctu = 1 << ctu_log2, cw = (width + ctu - 1) >> ctu_log2, ch = (height + ctu - 1) >> ctu_log2;
  const int n = cw * ch;
  int f[13];
  for (int cy = 0; cy < ch; cy++)
    for (int cx = 0; cx < cw; cx++) {
      const int idx = cy * cw + cx, x0 = cx << ctu_log2, y0 = cy << ctu_log2;
      const int x1 = x0 + ctu > width ? width : x0 + ctu, y1 = y0 + ctu > height ? height : y0 + ctu;
      if (ctu_enable[idx]) {
        const int is7 = params->luma_filter_7x7 != 0;
        for (int by = y0; by < y1; by += 4)
          for (int bx = x0; bx < x1; bx += 4) {
            const uint8_t cl = alf_classify_block(src[0], sstride[0], width, height, bx, by, bd_luma + 4);
            alf_transpose_coeff(params->luma_coeff[cl & 31], cl >> 5, is7, f);
            for (int y = by; y < by + 4; y++)
              for (int x = bx; x < bx + 4; x++)
                dst[0][(ptrdiff_t)y * dstride[0] + x] =
                    (Pel)alf_sample(src[0], sstride[0], width, height, x, y, f, is7, (1 << bd_luma) - 1);
          }
      }
      for (int c = 1; c < 3; c++) {
        if (!ctu_enable[c * n + idx]) continue;
        for (int i = 0; i < 7; i++) f[i] = params->chroma_coeff[i];
        for (int y = y0 >> 1; y < (y1 >> 1); y++)
          for (int x = x0 >> 1; x < (x1 >> 1); x++)
            dst[c][(ptrdiff_t)y * dstride[c] + x] =
                (Pel)alf_sample(src[c], sstride[c], width >> 1, height >> 1, x, y, f, 0, (1 << bd_chroma) - 1);
      }
    }
  return 0;
}

/* ------------------------------------------------------------------------------------------------
 * Encoder SAO statistics: EncSampleAdaptiveOffset::getStatistics
 * (EncoderLib/EncSampleAdaptiveOffset.cpp:278-331) + getBlkStats (:1122-1487), the path the encoder takes
 * without SaoCtuBoundary (isCalculatePreDeblockSamples == false): per CTU, component and SAO type, the count
 * of samples per class and the sum of (original - deblocked) per class.  The region of a CTU block that
 * contributes depends on the type, on the neighbour availability and on the "skip lines" of createEncData
 * (:122-128: right 5 / bottom 4 luma, 3 / 2 chroma -- applied when the right / below CTU exists).
 * out[ctu][comp][type][0..31] = diff, [32..63] = count (the memory layout of SAOStatData, EncSampleAdaptiveOffset.h:53-57).
 * avail[ctu]: ILF_AVAIL_L / _A / _AL of deriveLoopFilterBoundaryAvailibility (:1489-); right / below /
 * above-right come from the picture bounds, as in :307-309.
 * ---------------------------------------------------------------------------------------------- */
static void sao_stats_block(const Pel* src, ptrdiff_t ss, const Pel* org, ptrdiff_t os, int w, int h, int bd, int skip_r, int skip_b, int L, int R, int A,
                            int B, int AL, int AR, int64_t* out /* [5][64] */) {
  memset(out, 0, sizeof(int64_t) * 5 * 64);
  for (int type = 0; type < 5; type++) {
    int64_t* diff = out + type * 64;
    int64_t* count = diff + 32;
    int start_x, end_x, start_y, end_y;
    switch (type) {
      case ILF_SAO_EO_0: /* :1146-1163 */
        start_x = L ? 0 : 1; end_x = R ? w - skip_r : w - 1; end_y = B ? h - skip_b : h;
        for (int y = 0; y < end_y; y++)
          for (int x = start_x; x < end_x; x++) {
            const Pel* s = src + y * ss + x;
            const int c = 2 + sgn(s[0] - s[-1]) + sgn(s[0] - s[1]);
            diff[c] += org[y * os + x] - s[0]; count[c]++;
          }
        break;
      case ILF_SAO_EO_90: /* :1192-1230 */
        start_x = 0; end_x = R ? w - skip_r : w; start_y = A ? 0 : 1; end_y = B ? h - skip_b : h - 1;
        for (int y = start_y; y < end_y; y++)
          for (int x = start_x; x < end_x; x++) {
            const Pel* s = src + y * ss + x;
            const int c = 2 + sgn(s[0] - s[-ss]) + sgn(s[0] - s[ss]);
            diff[c] += org[y * os + x] - s[0]; count[c]++;
          }
        break;
      case ILF_SAO_EO_135: { /* :1258-1312: the first line has its own column range */
        start_x = L ? 0 : 1; end_x = R ? w - skip_r : w - 1; end_y = B ? h - skip_b : h - 1;
        const int fs = AL ? 0 : 1, fe = A ? end_x : 1;
        for (int y = 0; y < end_y; y++) {
          const int x0 = y == 0 ? fs : start_x, x1 = y == 0 ? fe : end_x;
          for (int x = x0; x < x1; x++) {
            const Pel* s = src + y * ss + x;
            const int c = 2 + sgn(s[0] - s[-ss - 1]) + sgn(s[0] - s[ss + 1]);
            diff[c] += org[y * os + x] - s[0]; count[c]++;
          }
        }
        break;
      }
      case ILF_SAO_EO_45: { /* :1340-1395 */
        start_x = L ? 0 : 1; end_x = R ? w - skip_r : w - 1; end_y = B ? h - skip_b : h - 1;
        const int fs = A ? start_x : end_x, fe = (!R && AR) ? w : end_x;
        for (int y = 0; y < end_y; y++) {
          const int x0 = y == 0 ? fs : start_x, x1 = y == 0 ? fe : end_x;
          for (int x = x0; x < x1; x++) {
            const Pel* s = src + y * ss + x;
            const int c = 2 + sgn(s[0] - s[-ss + 1]) + sgn(s[0] - s[ss - 1]);
            diff[c] += org[y * os + x] - s[0]; count[c]++;
          }
        }
        break;
      }
      default: /* ILF_SAO_BO :1437-1456 */
        end_x = R ? w - skip_r : w; end_y = B ? h - skip_b : h;
        for (int y = 0; y < end_y; y++)
          for (int x = 0; x < end_x; x++) {
            const int c = src[y * ss + x] >> (bd - 5);
            diff[c] += org[y * os + x] - src[y * ss + x]; count[c]++;
          }
        break;
    }
  }
}

int ilf_oracle_sao_stats(const int16_t* const rec[3], const ptrdiff_t rec_stride[3], const int16_t* const org[3], const ptrdiff_t org_stride[3], int width,
                         int height, int bd_luma, int bd_chroma, int ctu_log2, const uint8_t* avail, int64_t* out) {
  const int ctu = 1 << ctu_log2, cw = (width + ctu - 1) / ctu, ch = (height + ctu - 1) / ctu;
  for (int cy = 0; cy < ch; cy++)
    for (int cx = 0; cx < cw; cx++) {
      const int a = avail[cy * cw + cx];
      const int x0 = cx * ctu, y0 = cy * ctu;
      const int R = x0 + ctu < width, B = y0 + ctu < height, AR = y0 > 0 && R; /* :307-309 */
      for (int comp = 0; comp < 3; comp++) {
        const int sh = comp ? 1 : 0;
        const int bw = ((x0 + ctu > width ? width - x0 : ctu)) >> sh, bh = ((y0 + ctu > height ? height - y0 : ctu)) >> sh;
        sao_stats_block(rec[comp] + (ptrdiff_t)(y0 >> sh) * rec_stride[comp] + (x0 >> sh), rec_stride[comp],
                        org[comp] + (ptrdiff_t)(y0 >> sh) * org_stride[comp] + (x0 >> sh), org_stride[comp], bw, bh, comp ? bd_chroma : bd_luma, comp ? 3 : 5,
                        comp ? 2 : 4, (a & ILF_AVAIL_L) != 0, R, (a & ILF_AVAIL_A) != 0, B, (a & ILF_AVAIL_AL) != 0, AR,
                        out + ((size_t)(cy * cw + cx) * 3 + comp) * 5 * 64);
      }
    }
  return 0;
}

/* One block with explicit flags, for the unit comparison with the reference's getBlkStats (tests/test_oracle_units.py). */
int ilf_oracle_sao_stats_block(const int16_t* src, ptrdiff_t ss, const int16_t* org, ptrdiff_t os, int w, int h, int bd, int is_chroma, unsigned avail6,
                               int64_t* out) {
  sao_stats_block(src, ss, org, os, w, h, bd, is_chroma ? 3 : 5, is_chroma ? 2 : 4, avail6 & 1, (avail6 >> 1) & 1, (avail6 >> 2) & 1, (avail6 >> 3) & 1,
                  (avail6 >> 4) & 1, (avail6 >> 5) & 1, out);
  return 0;
}

/* ---------------------------------------------------------------------------------------------------------
 * Post-filter consumers: decoded-picture hash CRC / checksum and reference border extension.
 * --------------------------------------------------------------------------------------------------------- */
/* compCRC, PicYuvMD5.cpp:91-128: CRC-16 (poly 0x1021, register starts at 0xffff) over the low byte then -- above 8 bit -- the high
 * byte of every sample in raster order, message bits shifted in at the bottom, 16 zero bits appended. */
unsigned ilf_oracle_crc(const int16_t* plane, ptrdiff_t stride, int width, int height, int bit_depth) {
  unsigned crc = 0xffff;
  for (int y = 0; y < height; y++)
    for (int x = 0; x < width; x++) {
      const unsigned v = (uint16_t)plane[y * stride + x];
      for (int b = 0; b < (bit_depth > 8 ? 16 : 8); b++) {
        const unsigned msb = (crc >> 15) & 1, bit = b < 8 ? (v >> (7 - b)) & 1 : (v >> (23 - b)) & 1;
        crc = (((crc << 1) + bit) & 0xffff) ^ (msb * 0x1021);
      }
    }
  for (int b = 0; b < 16; b++) {
    const unsigned msb = (crc >> 15) & 1;
    crc = ((crc << 1) & 0xffff) ^ (msb * 0x1021);
  }
  return crc;
}

/* compChecksum, PicYuvMD5.cpp:144-163 */
unsigned ilf_oracle_checksum(const int16_t* plane, ptrdiff_t stride, int width, int height, int bit_depth) {
  uint32_t sum = 0;
  for (int y = 0; y < height; y++)
    for (int x = 0; x < width; x++) {
      const unsigned mask = ((x & 0xff) ^ (y & 0xff) ^ (x >> 8) ^ (y >> 8)) & 0xff;
      const unsigned v = (uint16_t)plane[y * stride + x];
      sum += (v & 0xff) ^ mask;
      if (bit_depth > 8) sum += (v >> 8) ^ mask;
    }
  return sum;
}

/* Picture::extendPicBorder, Picture.cpp:996-1040, one plane: `plane` points at sample (0, 0) of a buffer that has xmargin columns
 * and ymargin rows of room on every side. */
void ilf_oracle_extend_border(int16_t* plane, ptrdiff_t stride, int width, int height, int xmargin, int ymargin) {
  for (int y = 0; y < height; y++)
    for (int x = 0; x < xmargin; x++) {
      plane[y * stride - xmargin + x] = plane[y * stride];
      plane[y * stride + width + x] = plane[y * stride + width - 1];
    }
  for (int y = 0; y < ymargin; y++) {
    memcpy(plane + (height + y) * stride - xmargin, plane + (height - 1) * stride - xmargin, sizeof(int16_t) * (size_t)(width + 2 * xmargin));
    memcpy(plane - (y + 1) * stride - xmargin, plane - xmargin, sizeof(int16_t) * (size_t)(width + 2 * xmargin));
  }
}

/* ---------------------------------------------------------------------------------------------------------
 * Encoder ALF statistics: EncAdaptiveLoopFilter::deriveStatsForFiltering / getBlkStats / calcCovariance
 * (EncoderLib/EncAdaptiveLoopFilter.cpp:1317-1514).  Per CTU and class the covariance of the 13 (7) symmetric tap sums of the
 * reconstruction around every sample, their correlation with (original - reconstruction) and its energy.  The reference keeps
 * doubles that hold exact integers (< 2^53); here int64.  Layout per CTU (ILF_ALF_STATS_WORDS): luma [25 classes][105] with the
 * 7x7 shape (the 5x5 shape's statistics are the rows / columns {2, 5, 6, 7, 10, 11, 12} of it), then Cb [36], Cr [36] with the
 * 5x5 shape; one record = E upper triangle row-major (k <= l), then y[k], then pixAcc.
 * --------------------------------------------------------------------------------------------------------- */
/* tap k of the canonical (transposeIdx 0) enumeration of calcCovariance :1442-1461: rows -half..-1 left to right, then the left
 * half of the centre row; transposeIdx t maps the offset (a, b) to 0: (a, b), 1: (b, a), 2: (-a, b), 3: (b, -a) (:1463-1510) */
static void alf_stats_taps(int half, int t, int dx[12], int dy[12]) {
  int k = 0;
  for (int b = -half; b <= 0; b++)
    for (int a = -(half + b); a <= (b < 0 ? half + b : -1); a++, k++) {
      dx[k] = t == 0 ? a : (t == 1 ? b : (t == 2 ? -a : b));
      dy[k] = t == 0 ? b : (t == 1 ? a : (t == 2 ? b : -a));
    }
}

/* rec: picture padded by replication (clamped reads); cls: classIdx | transposeIdx << 5 per 4x4 block (NULL: class 0, no transpose) */
static void alf_stats_plane(const Pel* rec, ptrdiff_t rs, const Pel* org, ptrdiff_t os, int w, int h, int half, const uint8_t* cls, int units_w, int ctu_log2_c,
                            int ctus_w, long long* out, size_t ctu_words, size_t plane_off, int rec_words) {
  const int n = half == 3 ? 13 : 7;
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) {
      const int c = cls ? cls[(y >> 2) * units_w + (x >> 2)] : 0;
      int dx[12], dy[12], e[13];
      alf_stats_taps(half, c >> 5, dx, dy);
      for (int k = 0; k < n - 1; k++) e[k] = px(rec, rs, w, h, x + dx[k], y + dy[k]) + px(rec, rs, w, h, x - dx[k], y - dy[k]);
      e[n - 1] = rec[y * rs + x];
      const int yl = org[y * os + x] - rec[y * rs + x];
      long long* o = out + (size_t)((y >> ctu_log2_c) * ctus_w + (x >> ctu_log2_c)) * ctu_words + plane_off + (size_t)(c & 31) * rec_words;
      int i = 0;
      for (int k = 0; k < n; k++)
        for (int l = k; l < n; l++) o[i++] += (long long)e[k] * e[l];
      for (int k = 0; k < n; k++) o[i++] += (long long)e[k] * yl;
      o[i] += (long long)yl * yl;
    }
}

int ilf_oracle_alf_stats(const int16_t* const rec[3], const ptrdiff_t rec_stride[3], const int16_t* const org[3], const ptrdiff_t org_stride[3], int width, int height,
                         int bd_luma, int ctu_log2, long long* out) {
  const int uw = width / 4, uh = height / 4, ctus_w = (width + (1 << ctu_log2) - 1) >> ctu_log2, ctus_h = (height + (1 << ctu_log2) - 1) >> ctu_log2;
  const size_t words = 25 * 105 + 36 + 36;
  uint8_t* cls = (uint8_t*)malloc((size_t)uw * uh);
  if (!cls) return -1;
  ilf_oracle_alf_classify(rec[0], rec_stride[0], width, height, bd_luma, cls);
  memset(out, 0, sizeof(long long) * words * ctus_w * ctus_h);
  alf_stats_plane(rec[0], rec_stride[0], org[0], org_stride[0], width, height, 3, cls, uw, ctu_log2, ctus_w, out, words, 0, 105);
  alf_stats_plane(rec[1], rec_stride[1], org[1], org_stride[1], width / 2, height / 2, 2, NULL, 0, ctu_log2 - 1, ctus_w, out, words, 25 * 105, 36);
  alf_stats_plane(rec[2], rec_stride[2], org[2], org_stride[2], width / 2, height / 2, 2, NULL, 0, ctu_log2 - 1, ctus_w, out, words, 25 * 105 + 36, 36);
  free(cls);
  return 0;
}

/* one plane as ONE block with explicit classes (unit test against the reference's getBlkStats): out[classes][n (n + 1) / 2 + n + 1] */
int ilf_oracle_alf_stats_block(const int16_t* rec, ptrdiff_t rs, const int16_t* org, ptrdiff_t os, int w, int h, int shape7, const uint8_t* cls_per_unit, long long* out) {
  const int n = shape7 ? 13 : 7, words = n * (n + 1) / 2 + n + 1;
  memset(out, 0, sizeof(long long) * (size_t)words * (cls_per_unit ? 25 : 1));
  alf_stats_plane(rec, rs, org, os, w, h, shape7 ? 3 : 2, cls_per_unit, (w + 3) / 4, 30, 1, out, 0, 0, words);
  return 0;
}
