// Host check of the dot-product coefficient layout (csrc/ilf_alf_tab.cuh): entries built by build_entry and evaluated slot
// by slot with a scalar model of IDP.2A must equal the direct diamond sum of AdaptiveLoopFilter::filterBlk for every output
// parity, both filter shapes, the 5x5 luma filter embedded in the 7x7 layout, and coefficients at the limits of the path.
#define __host__
#define __device__
#define __forceinline__ inline
#include <cstdio>
#include <cstdlib>
#include <random>
#include "ilf_alf_tab.cuh"

using namespace ilf::alftab;

static int dp2a(uint32_t a, uint32_t b, int half, int c) {
  const int a0 = a & 0xFFFF, a1 = a >> 16;
  const int b0 = (int8_t)(b >> (16 * half)), b1 = (int8_t)(b >> (16 * half + 8));
  return c + a0 * b0 + a1 * b1;
}

template <int R, int RT>
static int run(std::mt19937& rng, int nwords, int cases) {
  constexpr int NC = RT * (RT + 1) * 2 + 1 - RT * (RT + 1);  // (RT+1)^2 ... number of coefficients = RT(RT+1)+1
  int bad = 0;
  for (int it = 0; it < cases; it++) {
    int f[16] = {0};
    const int n = RT * (RT + 1) + 1;
    for (int k = 0; k < n - 1; k++) {
      const int r = rng() % 8;
      f[k] = r == 0 ? -128 : (r == 1 ? 127 : (int)(rng() % 256) - 128);
    }
    // the taps next to the centre and the centre may be large
    f[coef_index<RT>(1, 0)] = (int)(rng() % 3001) - 1500;
    f[coef_index<RT>(0, 1)] = (int)(rng() % 3001) - 1500;
    int sum = 0;
    for (int k = 0; k < n - 1; k++) sum += f[k];
    f[n - 1] = it % 3 == 0 ? 512 - 2 * sum : (int)(rng() % 32001) - 16000;
    uint32_t e[32];
    if (!build_entry<R, RT>(f, e, nwords, nullptr)) { printf("entry rejected unexpectedly\n"); bad++; continue; }
    // window of 16 x 16 samples, outputs at (8, 8) and (9, 8)
    int smp[16][16];
    for (auto& row : smp) for (int& v : row) v = rng() % 4096;
    for (int p = 0; p < 2; p++) {
      const int x = 8 + p, y = 8;
      long long want = 0;
      for (int dy = -RT; dy <= RT; dy++)
        for (int dx = -RT; dx <= RT; dx++) { const int k = coef_index<RT>(dx, dy); if (k >= 0) want += (long long)f[k] * smp[y + dy][x + dx]; }
      int acc = 0, hi = 0;
      const int nr = num_regs<R>();
      for (int dy = -RT; dy <= RT; dy++)
        for (int q = 0; q <= R; q++) {
          if (!holds<R, RT>(p, dy, q)) continue;
          const int s = slot<R>(p, dy, q), xs = x + dx0<R>(p, q);
          if (xs & 1) { printf("slot not word aligned\n"); bad++; }
          acc = dp2a((uint32_t)smp[y + dy][xs] | (uint32_t)smp[y + dy][xs + 1] << 16, e[p * nr + (s >> 1)], s & 1, acc);
        }
      for (int dy = -1; dy <= 1; dy++)
        for (int q = 0; q <= 1; q++) {
          if (!holds<1, 1>(p, dy, q)) continue;
          const int s = slot<1>(p, dy, q), xs = x + dx0<1>(p, q);
          if (xs & 1) { printf("high slot not word aligned\n"); bad++; }
          hi = dp2a((uint32_t)smp[y + dy][xs] | (uint32_t)smp[y + dy][xs + 1] << 16, e[2 * nr + p * 2 + (s >> 1)], s & 1, hi);
        }
      const long long got = (long long)acc + ((long long)hi << HI_SHIFT);
      if (got != want) { if (bad < 5) printf("R=%d RT=%d parity %d: got %lld want %lld\n", R, RT, p, got, want); bad++; }
    }
    // slots of the layout that hold no tap of the filter must be zero (the kernel skips them)
    for (int p = 0; p < 2; p++)
      for (int dy = -R; dy <= R; dy++)
        for (int q = 0; q <= R; q++)
          if (holds<R, R>(p, dy, q) && !holds<R, RT>(p, dy, q)) {
            const int s = slot<R>(p, dy, q);
            if ((e[p * num_regs<R>() + (s >> 1)] >> (16 * (s & 1))) & 0xFFFF) { printf("skipped slot not empty\n"); bad++; }
          }
  }
  (void)NC;
  return bad;
}

int main() {
  std::mt19937 rng(7);
  int bad = 0;
  bad += run<3, 3>(rng, LUMA_WORDS, 2000);
  bad += run<3, 2>(rng, LUMA_WORDS, 2000);
  bad += run<2, 2>(rng, CHROMA_WORDS, 2000);
  // every slot index is used exactly once per parity
  for (int p = 0; p < 2; p++) {
    int seen[16] = {0};
    for (int dy = -3; dy <= 3; dy++) for (int q = 0; q <= 3; q++) if (holds<3, 3>(p, dy, q)) seen[slot<3>(p, dy, q)]++;
    for (int i = 0; i < 16; i++) if (seen[i] != 1) { printf("7x7 slot %d used %d times\n", i, seen[i]); bad++; }
  }
  // out-of-range coefficients are rejected
  { int f[16] = {0}; f[0] = 128; f[12] = 512; uint32_t e[32]; if (build_entry<3, 3>(f, e, LUMA_WORDS, nullptr)) { printf("128 accepted\n"); bad++; } }
  { int f[16] = {0}; f[0] = -129; f[12] = 512; uint32_t e[32]; if (build_entry<3, 3>(f, e, LUMA_WORDS, nullptr)) { printf("-129 accepted\n"); bad++; } }
  { int f[16] = {0}; f[12] = 16400; uint32_t e[32]; if (build_entry<3, 3>(f, e, LUMA_WORDS, nullptr)) { printf("centre 16400 accepted\n"); bad++; } }
  printf(bad ? "FAILED: %d\n" : "ok\n", bad);
  return bad != 0;
}
