// pipes2.cu -- which issue pipe does each integer instruction use?  Pairs every instruction X with IMAD (fma pipe) and with
// LOP3 (alu pipe) 1:1 on independent chains: the pairing that runs faster than X alone shares no pipe with X.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
constexpr int ITER = 512, CH = 8, UNR = 8;
enum { IMAD, LOP3, IADD3, VADD2, VSUB2, VMAXS2, VMINU2, VIADDMAX, PRMT, SHF, IABS, DP2A, SHL, ISETP_SEL, VIMNMX, LEA, FFMA, FFMA2, FADDRM, I2F, F2I, FMNMX, HFMA2, IMADHI, NKIND };
const char* names[NKIND] = {"IMAD", "LOP3", "IADD3", "VIADD.16x2(add)", "VIADD.16x2(sub)", "VIMNMX.S16x2", "VIMNMX.U16x2", "VIADDMNMX.16x2", "PRMT", "SHF", "IABS", "IDP.2A", "SHL(imm)", "ISETP+SEL", "VIMNMX(32)", "LEA", "FFMA", "FFMA2(f32x2)", "FADD.RM", "I2F", "F2I", "FMNMX", "HFMA2", "IMAD.SHL/big"};
template <int K>
__device__ __forceinline__ void op(uint32_t& x, uint32_t a, uint32_t b) {
  if (K == IMAD) x = x * a + b;
  else if (K == LOP3) x = (x & a) ^ b;
  else if (K == IADD3) asm volatile("add.u32 %0, %0, %1;" : "+r"(x) : "r"(a));
  else if (K == VADD2) x = __vadd2(x, a);
  else if (K == VSUB2) x = __vsub2(x, a);
  else if (K == VMAXS2) x = __vmaxs2(x, a);
  else if (K == VMINU2) x = __vminu2(x, a);
  else if (K == VIADDMAX) x = __viaddmax_s16x2(x, a, b);
  else if (K == PRMT) x = __byte_perm(x, a, b);
  else if (K == SHF) x = __funnelshift_r(x, a, 16);
  else if (K == IABS) x = (uint32_t)abs((int)x);
  else if (K == DP2A) asm volatile("dp2a.lo.s32.s32 %0, %0, %1, %2;" : "+r"(x) : "r"(a), "r"(b));
  else if (K == SHL) x = (x << 3);
  else if (K == ISETP_SEL) x = ((int)x > (int)a) ? b : x;
  else if (K == VIMNMX) x = (uint32_t)max((int)x, (int)a);
  else if (K == LEA) x = (x << 2) + a;
  else if (K == FFMA2) {
    uint32_t y = x ^ 0x3f800000u;
    asm volatile("{ .reg .b64 t, u, v; mov.b64 t, {%0, %1}; mov.b64 u, {%2, %2}; mov.b64 v, {%3, %3}; fma.rn.f32x2 t, t, u, v; mov.b64 {%0, %1}, t; }" : "+r"(x), "+r"(y) : "r"(a), "r"(b));
    x ^= y & 1;
  }
  else if (K == FFMA) x = __float_as_uint(fmaf(__uint_as_float(x), __uint_as_float(a), __uint_as_float(b)));
  else if (K == FADDRM) x = __float_as_uint(__fadd_rd(__uint_as_float(x), __uint_as_float(a)));
  else if (K == I2F) x = __float_as_uint((float)(int)x) + a;
  else if (K == F2I) x = (uint32_t)__float2int_rd(__uint_as_float(x | 0x40000000u));
  else if (K == FMNMX) x = __float_as_uint(fminf(__uint_as_float(x), __uint_as_float(a)));
  else if (K == HFMA2) asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(x) : "r"(a), "r"(b));
  else if (K == IMADHI) x = x * 0x10001u + b;
}
template <int K0, int K1>
__global__ void __launch_bounds__(512) bench(uint32_t* out, uint32_t a, uint32_t b, long long* cyc) {
  uint32_t x[CH];
#pragma unroll
  for (int i = 0; i < CH; i++) x[i] = threadIdx.x * 7 + i;
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITER; it++) {
#pragma unroll
    for (int u = 0; u < UNR; u++)
#pragma unroll
      for (int i = 0; i < CH; i++) { if (i & 1) op<K1>(x[i], a, b); else op<K0>(x[i], a, b); }
  }
  __syncthreads();
  const long long t1 = clock64();
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < CH; i++) s ^= x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
static uint32_t* out; static long long* cyc; static int blocks;
template <int K0, int K1> double run() {
  bench<K0, K1><<<blocks, 512>>>(out, 3, 0x5410, cyc);
  bench<K0, K1><<<blocks, 512>>>(out, 3, 0x5410, cyc);
  cudaDeviceSynchronize();
  static long long h[4096];
  cudaMemcpy(h, cyc, blocks * 8, cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < blocks; i++) avg += h[i]; avg /= blocks;
  return 16.0 * ITER * CH * UNR / avg;
}
template <int K> void row() {
  const double alone = run<K, K>(), with_fma = run<K, IMAD>(), with_alu = run<K, LOP3>();
  printf("%-18s alone %.3f   +IMAD %.3f   +LOP3 %.3f   -> %s\n", names[K], alone, with_fma, with_alu,
         with_fma > 1.25 * alone && with_alu < 1.25 * alone ? "ALU pipe" : (with_alu > 1.25 * alone && with_fma < 1.25 * alone ? "FMA pipe" : "both/unclear"));
}
int main() {
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  blocks = sms * 4;
  cudaMalloc(&out, blocks * 512 * 4); cudaMalloc(&cyc, blocks * 8);
  printf("warp-instructions per clock64 tick per SMSP (16 resident warps per SMSP)\n");
  row<IMAD>(); row<LOP3>(); row<IADD3>(); row<VADD2>(); row<VSUB2>(); row<VMAXS2>(); row<VMINU2>(); row<VIADDMAX>(); row<PRMT>(); row<SHF>(); row<IABS>(); row<DP2A>();
  row<SHL>(); row<ISETP_SEL>(); row<VIMNMX>(); row<LEA>(); row<FFMA>(); row<FFMA2>(); row<FADDRM>(); row<I2F>(); row<F2I>(); row<FMNMX>(); row<HFMA2>(); row<IMADHI>();
  return 0;
}
