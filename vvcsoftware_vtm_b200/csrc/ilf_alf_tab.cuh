// ilf_alf_tab.cuh -- coefficient layout of the ALF "dot-product" filter path (shared by the kernels and by ilf_set_alf_params).
//
// The diamond filters of AdaptiveLoopFilter::filterBlk (AdaptiveLoopFilter.cpp:465-650) multiply 16-bit samples by small
// coefficients.  sm_100a has IDP.2A: d = c + a.lo16 * b.byte[k] + a.hi16 * b.byte[k + 1] (k = 0: .LO, k = 2: .HI) -- two taps per
// instruction, straight from the packed int16 words of a row as they lie in memory, with no unpacking and no pair sums.
// For an output sample at x the words of row y + dy that start at an even (x even) or odd (x odd) offset from x are the
// memory-aligned ones; each such word is a "slot" with two coefficient bytes.  A diamond of radius R has (R + 1)^2 slots per
// output parity.
//
// Coefficients are int16 in the bitstream; the ones next to the centre and the centre itself are often outside int8
// (centre ~ 150..1000 of 512 = 1.0).  A tap with |dx| + |dy| <= 1 outside int8 is therefore split c = lo + 128 hi, lo in [-64, 63],
// and the high parts form a second, radius-1 diamond (4 slots per parity) whose sum is shifted left by 7.  All other taps
// must fit int8 -- true for the filters real encoders produce at >= 1080p, not guaranteed; ilf_set_alf_params checks
// the picture's filters and selects the general 32-bit multiply path when one does not fit (bit-exact either way).
#pragma once
#include <stdint.h>

namespace ilf {
namespace alftab {

#define ALFTAB_FN __host__ __device__ __forceinline__ constexpr

ALFTAB_FN int iabs(int v) { return v < 0 ? -v : v; }
// offset (from the output sample) of the first sample of word slot q, for output parity p (x & 1)
template <int R> ALFTAB_FN int dx0(int p, int q) { return (p ? -(2 * (R / 2) + 1) : -2 * ((R + 1) / 2)) + 2 * q; }
// does slot (p, dy, q) of a radius-R layout hold a tap of the radius-RT diamond?
template <int R, int RT> ALFTAB_FN bool holds(int p, int dy, int q) {
  return iabs(dy) <= RT && (iabs(dx0<R>(p, q)) <= RT - iabs(dy) || iabs(dx0<R>(p, q) + 1) <= RT - iabs(dy));
}
template <int R> ALFTAB_FN int first_q(int p, int dy) {
  int q = 0;
  while (q < R && !holds<R, R>(p, dy, q)) q++;
  return q;
}
// slots are numbered row by row (dy = -R .. R), left to right; slot s lives in register s >> 1, half s & 1
template <int R> ALFTAB_FN int slot(int p, int dy, int q) {
  const int n = R + 1 - iabs(dy);
  return (dy <= 0 ? n * (n - 1) / 2 : (R + 1) * (R + 1) - n * (n + 1) / 2) + q - first_q<R>(p, dy);
}
template <int R> ALFTAB_FN int num_slots() { return (R + 1) * (R + 1); }
template <int R> ALFTAB_FN int num_regs() { return ((R + 1) * (R + 1) + 1) / 2; }

static_assert(slot<3>(0, -3, 2) == 0 && slot<3>(0, 0, 0) == 6 && slot<3>(0, 3, 2) == 15 && slot<3>(1, 3, 1) == 15 && slot<3>(1, 1, 0) == 10, "7x7 slot numbering");
static_assert(slot<2>(0, -2, 1) == 0 && slot<2>(1, 0, 0) == 3 && slot<2>(0, 2, 1) == 8, "5x5 slot numbering");
static_assert(slot<1>(0, -1, 1) == 0 && slot<1>(0, 0, 0) == 1 && slot<1>(1, 0, 1) == 2 && slot<1>(1, 1, 0) == 3, "high-part slot numbering");

// Index of the coefficient of tap (dx, dy) in the reference's coefficient order of a radius-RT diamond
// (AdaptiveLoopFilter.cpp:577-650: rows from the top of the diamond down to the centre row, centre last), or -1 outside.
template <int RT> ALFTAB_FN int coef_index(int dx, int dy) {
  if (iabs(dx) + iabs(dy) > RT) return -1;
  if (dy < 0 || (dy == 0 && dx < 0)) { dx = -dx; dy = -dy; }
  const int m = RT - dy;
  return m * (m + 1) - dx;
}
static_assert(coef_index<3>(0, 0) == 12 && coef_index<3>(1, 0) == 11 && coef_index<3>(-3, 0) == 9 && coef_index<3>(2, 1) == 4 && coef_index<3>(-2, 1) == 8 &&
              coef_index<3>(2, -1) == 8 && coef_index<3>(1, 2) == 1 && coef_index<3>(0, -3) == 0 && coef_index<2>(0, 0) == 6 && coef_index<2>(-1, 1) == 3, "coefficient order");

// Words of one table entry.  Luma: radius-3 layout (a 5x5 luma filter sits in it at the same sample positions):
// [0, 8) even-x slots, [8, 16) odd-x slots, [16, 18) / [18, 20) high parts.  Chroma: radius-2 layout: [0, 5), [5, 10), [10, 12), [12, 14).
constexpr int LUMA_WORDS = 20, CHROMA_WORDS = 16 /* 14 used */;
constexpr int HI_SHIFT = 7;

// Host side: builds one entry from the coefficients f[] (reference order of the radius-RT diamond) in a radius-R layout.
// Returns false when a coefficient does not fit the layout (the caller then uses the general path); *neigh_hi is set when one of
// the four neighbours of the centre needed a high part (else only the centre has one and the kernels skip three of the four
// high-part slots).
template <int R, int RT>
inline bool build_entry(const int* f, uint32_t* out, int nwords, bool* neigh_hi) {
  for (int i = 0; i < nwords; i++) out[i] = 0;
  bool ok = true;
  auto hi_of = [](int c) { return c >= -128 && c <= 127 ? 0 : (c + 64) >> HI_SHIFT; };
  auto lo_part = [&](int dx, int dy) -> int {
    const int k = coef_index<RT>(dx, dy);
    if (k < 0) return 0;
    const int c = f[k];
    if (iabs(dx) + iabs(dy) <= 1) return c - (hi_of(c) << HI_SHIFT);
    if (c < -128 || c > 127) ok = false;
    return c;
  };
  auto hi_part = [&](int dx, int dy) -> int {
    const int k = coef_index<RT>(dx, dy);
    if (k < 0 || iabs(dx) + iabs(dy) > 1) return 0;
    const int h = hi_of(f[k]);
    if (h < -128 || h > 127) ok = false;
    if (h && (dx || dy) && neigh_hi) *neigh_hi = true;
    return h;
  };
  const int nr = num_regs<R>();
  for (int p = 0; p < 2; p++) {
    for (int dy = -R; dy <= R; dy++)
      for (int q = 0; q <= R; q++) {
        if (!holds<R, R>(p, dy, q)) continue;
        const int s = slot<R>(p, dy, q), x = dx0<R>(p, q);
        const uint32_t pair = (uint32_t)(uint8_t)lo_part(x, dy) | (uint32_t)(uint8_t)lo_part(x + 1, dy) << 8;
        out[p * nr + (s >> 1)] |= pair << (16 * (s & 1));
      }
    for (int dy = -1; dy <= 1; dy++)
      for (int q = 0; q <= 1; q++) {
        if (!holds<1, 1>(p, dy, q)) continue;
        const int s = slot<1>(p, dy, q), x = dx0<1>(p, q);
        const uint32_t pair = (uint32_t)(uint8_t)hi_part(x, dy) | (uint32_t)(uint8_t)hi_part(x + 1, dy) << 8;
        out[2 * nr + p * 2 + (s >> 1)] |= pair << (16 * (s & 1));
      }
  }
  return ok;
}

}  // namespace alftab
}  // namespace ilf
