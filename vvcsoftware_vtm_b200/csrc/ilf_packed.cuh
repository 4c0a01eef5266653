// ilf_packed.cuh -- two int16 samples per 32-bit register: the per-lane integer instructions sm_100a has
// (VIADD.16x2, VIMNMX.{S,U}16x2, VIMNMX3, VIADDMNMX.S16x2[.RELU], PRMT) wrapped so that the same arithmetic also
// compiles as plain host C++ (tests/packed_host_test.cu checks every helper against scalar code on the CPU).
// "Lane" = one 16-bit half of the register; lane 0 = bits 0..15 = the sample with the lower x coordinate.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define ILF_PK __host__ __device__ __forceinline__
#else
#define ILF_PK inline
#endif

namespace ilf {
namespace pk {

ILF_PK uint32_t pack(int lo, int hi) { return (uint32_t)(uint16_t)lo | ((uint32_t)(uint16_t)hi << 16); }
ILF_PK uint32_t splat(int v) { return pack(v, v); }
ILF_PK int lane0(uint32_t v) { return (int)(int16_t)(v & 0xFFFF); }
ILF_PK int lane1(uint32_t v) { return (int)(int16_t)(v >> 16); }

#if defined(__CUDA_ARCH__)
ILF_PK uint32_t add2(uint32_t a, uint32_t b) { return __vadd2(a, b); }
ILF_PK uint32_t sub2(uint32_t a, uint32_t b) { return __vsub2(a, b); }
ILF_PK uint32_t mins2(uint32_t a, uint32_t b) { return __vmins2(a, b); }
ILF_PK uint32_t maxs2(uint32_t a, uint32_t b) { return __vmaxs2(a, b); }
ILF_PK uint32_t minu2(uint32_t a, uint32_t b) { return __vminu2(a, b); }
ILF_PK uint32_t addmin2(uint32_t a, uint32_t b, uint32_t c) { return __viaddmin_s16x2(a, b, c); }            // min(a + b, c)
ILF_PK uint32_t addmax2(uint32_t a, uint32_t b, uint32_t c) { return __viaddmax_s16x2(a, b, c); }            // max(a + b, c)
ILF_PK uint32_t addmin2_relu(uint32_t a, uint32_t b, uint32_t c) { return __viaddmin_s16x2_relu(a, b, c); }  // max(min(a + b, c), 0)
ILF_PK uint32_t min2_relu(uint32_t a, uint32_t b) { return __vimin_s16x2_relu(a, b); }                       // max(min(a, b), 0)
// PRMT with the full 4-bit selectors (bit 3 = replicate the sign of the selected byte); __byte_perm masks that bit away.
ILF_PK uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
  return d;
}
ILF_PK uint32_t funnel16(uint32_t lo, uint32_t hi) { return __funnelshift_r(lo, hi, 16); }                  // [lane1(lo), lane0(hi)]
#else
#define ILF_PK_LANES(expr)                                   \
  int r0, r1;                                                \
  { const int x = lane0(a), y = lane0(b); (void)y; r0 = (expr); } \
  { const int x = lane1(a), y = lane1(b); (void)y; r1 = (expr); } \
  return pack(r0, r1)
inline int wrap16(int v) { return (int)(int16_t)(uint16_t)v; }
inline uint32_t add2(uint32_t a, uint32_t b) { ILF_PK_LANES(wrap16(x + y)); }
inline uint32_t sub2(uint32_t a, uint32_t b) { ILF_PK_LANES(wrap16(x - y)); }
inline uint32_t mins2(uint32_t a, uint32_t b) { ILF_PK_LANES(x < y ? x : y); }
inline uint32_t maxs2(uint32_t a, uint32_t b) { ILF_PK_LANES(x > y ? x : y); }
inline uint32_t minu2(uint32_t a, uint32_t b) { ILF_PK_LANES((uint16_t)x < (uint16_t)y ? x : y); }
inline uint32_t addmin2(uint32_t a, uint32_t b, uint32_t c) { return mins2(add2(a, b), c); }
inline uint32_t addmax2(uint32_t a, uint32_t b, uint32_t c) { return maxs2(add2(a, b), c); }
inline uint32_t addmin2_relu(uint32_t a, uint32_t b, uint32_t c) { return maxs2(mins2(add2(a, b), c), 0u); }
inline uint32_t min2_relu(uint32_t a, uint32_t b) { return maxs2(mins2(a, b), 0u); }
inline uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
  const uint64_t src = (uint64_t)a | ((uint64_t)b << 32);
  uint32_t out = 0;
  for (int i = 0; i < 4; i++) {
    const unsigned n = (sel >> (4 * i)) & 0xF;
    unsigned byte = (unsigned)(src >> (8 * (n & 7))) & 0xFF;
    if (n & 8) byte = (byte & 0x80) ? 0xFF : 0x00;
    out |= byte << (8 * i);
  }
  return out;
}
inline uint32_t funnel16(uint32_t lo, uint32_t hi) { return (lo >> 16) | (hi << 16); }
#undef ILF_PK_LANES
#endif

// ---------------------------------------------------------------------------------------------------------
// SAO (SampleAdaptiveOffset::offsetBlock, SampleAdaptiveOffset.cpp:292-508) on two samples at a time.
// ---------------------------------------------------------------------------------------------------------

// Edge-offset class index of two samples: sgn(c - a) + sgn(c - b) + 2 in each lane (0..4).  Samples are
// non-negative and below 2^13, so c + 2^13 - a never borrows across the lanes in a plain 32-bit add.
ILF_PK uint32_t sao_eo_index2(uint32_t c, uint32_t a, uint32_t b) {
  const uint32_t K = 0x20002000u;
  uint32_t ta = c + K - a, tb = c + K - b;                       // lanes in [1, 2^14 - 1]
  ta = mins2(maxs2(ta, 0x1FFF1FFFu), 0x20012001u);                // clamp to K-1 .. K+1
  tb = mins2(maxs2(tb, 0x1FFF1FFFu), 0x20012001u);
  return ta + tb - 0x3FFE3FFEu;                                   // (K-1)*2 subtracted -> 0..4
}

// Band-offset index of two samples: ((c >> shift) - band_pos) & 31, clamped to 4 (bands 0..3 carry offsets).
// nband = splat(32 - band_pos).
ILF_PK uint32_t sao_bo_index2(uint32_t c, int shift, uint32_t nband) {
  const uint32_t t = (c >> shift) & 0x001F001Fu;
  return minu2((t + nband) & 0x001F001Fu, 0x00040004u);
}

// out = clip(c + lut[idx], 0, max) per lane; lut = 5 signed bytes (lut_lo = bytes 0..3, lut_hi = byte 4), idx lanes 0..4.
ILF_PK uint32_t sao_apply2(uint32_t c, uint32_t idx, uint32_t lut_lo, uint32_t lut_hi, uint32_t maxv) {
  // selector nibbles: byte0 <- lut[idx0], byte1 <- sign(lut[idx0]), byte2 <- lut[idx1], byte3 <- sign(lut[idx1])
  // idx lanes hold 0..4 only: gather their low bytes with one PRMT (bytes 0 and 2 -> 0 and 1), then spread each index over its
  // two nibbles with a multiply (FMA pipe; the SAO kernel is ALU-pipe bound)
  const uint32_t sel = prmt(idx, 0u, 0x4420u) * 0x11u + 0x8080u;
  const uint32_t off = prmt(lut_lo, lut_hi, sel);
  return addmin2_relu(c, off, maxv);
}

}  // namespace pk
}  // namespace ilf
