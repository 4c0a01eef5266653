// Host check of vvcsoftware_vtm_b200/csrc/ilf_packed.cuh: every packed helper against plain scalar code.
// Built and run by tests/test_packed_math.py with g++ (no GPU needed).  Exit code 0 = all good.
#include <cstdio>
#include <cstdlib>
#include <random>

#include "ilf_packed.cuh"

using namespace ilf::pk;
static int fails = 0;
#define CHECK(cond, ...) do { if (!(cond)) { if (fails < 20) { printf(__VA_ARGS__); printf("\n"); } fails++; } } while (0)

static int sgn(int v) { return (v > 0) - (v < 0); }
static int clip(int lo, int hi, int v) { return v < lo ? lo : (v > hi ? hi : v); }

int main() {
  std::mt19937 rng(12345);
  for (int bd = 8; bd <= 12; bd++) {
    const int maxv = (1 << bd) - 1;
    for (int it = 0; it < 200000; it++) {
      auto rs = [&]() { const unsigned r = rng(); return (r & 7) == 0 ? ((r & 8) ? maxv : 0) : (int)((r >> 4) % (maxv + 1)); };
      const int c0 = rs(), c1 = rs(), a0 = (rng() & 3) ? rs() : c0, a1 = rs(), b0 = rs(), b1 = (rng() & 3) ? rs() : c1;
      int8_t lut[5];
      for (int k = 0; k < 5; k++) lut[k] = (int8_t)(rng() % 256);
      const uint32_t lut_lo = (uint8_t)lut[0] | ((uint32_t)(uint8_t)lut[1] << 8) | ((uint32_t)(uint8_t)lut[2] << 16) | ((uint32_t)(uint8_t)lut[3] << 24), lut_hi = (uint8_t)lut[4];
      // edge offset
      const uint32_t idx = sao_eo_index2(pack(c0, c1), pack(a0, a1), pack(b0, b1));
      const int e0 = sgn(c0 - a0) + sgn(c0 - b0) + 2, e1 = sgn(c1 - a1) + sgn(c1 - b1) + 2;
      CHECK(lane0(idx) == e0 && lane1(idx) == e1, "eo index bd%d c=%d,%d a=%d,%d b=%d,%d got %d,%d want %d,%d", bd, c0, c1, a0, a1, b0, b1, lane0(idx), lane1(idx), e0, e1);
      const uint32_t o = sao_apply2(pack(c0, c1), idx, lut_lo, lut_hi, splat(maxv));
      CHECK(lane0(o) == clip(0, maxv, c0 + lut[e0]) && lane1(o) == clip(0, maxv, c1 + lut[e1]), "eo apply bd%d", bd);
      // band offset
      const int band = rng() % 32;
      const uint32_t bi = sao_bo_index2(pack(c0, c1), bd - 5, splat((32 - band) & 31));
      const int k0 = ((c0 >> (bd - 5)) - band) & 31, k1 = ((c1 >> (bd - 5)) - band) & 31;
      CHECK(lane0(bi) == (k0 < 4 ? k0 : 4) && lane1(bi) == (k1 < 4 ? k1 : 4), "bo index bd%d c=%d,%d band %d got %d,%d", bd, c0, c1, band, lane0(bi), lane1(bi));
      int8_t bl[5] = {lut[0], lut[1], lut[2], lut[3], 0};
      const uint32_t ob = sao_apply2(pack(c0, c1), bi, lut_lo, 0, splat(maxv));
      CHECK(lane0(ob) == clip(0, maxv, c0 + bl[k0 < 4 ? k0 : 4]) && lane1(ob) == clip(0, maxv, c1 + bl[k1 < 4 ? k1 : 4]), "bo apply bd%d", bd);
    }
  }
#ifdef ILF_HAVE_DEBLOCK_PACKED
#include "packed_deblock_test.inc"
#endif
  printf("packed_host_test: %d failures\n", fails);
  return fails ? 1 : 0;
}
