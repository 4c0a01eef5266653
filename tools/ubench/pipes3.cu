// pipes3.cu -- issue rates of the instructions the filter kernels use, with chains the compiler cannot fold (every
// operation mixes in the neighbouring chain's value), and clock64 calibrated against globaltimer.  Prints warp-instructions
// per SM clock per SM sub-partition; rows "A+B" interleave the two kinds 1:1 and print the TOTAL rate (a total above
// either alone = the two issue to different pipes).
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o pipes3 pipes3.cu ; run on the GPU box.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
constexpr int ITER = 512, CH = 8, UNR = 8;
enum { IMAD, IADD3, LOP3, SHF, PRMT, VADD2, VMNMX2, VADDMNMX2, FFMA, FADD, FMNMX, I2FP, IDP2A, IMNMX, FMUL, HFMA2, LDS64, NK };
const char* names[NK] = {"IMAD", "IADD3", "LOP3", "SHF", "PRMT", "VIADD.16x2", "VIMNMX.16x2", "VIADDMNMX.16x2", "FFMA", "FADD", "FMNMX", "I2FP", "IDP.2A", "VIMNMX.S32", "FMUL", "HFMA2", "LDS.64"};
template <int K>
__device__ __forceinline__ void op(uint32_t& x, uint32_t y, uint32_t a, const uint32_t* sm) {
  if (K == IMAD) x = x * a + y;
  else if (K == IADD3) asm volatile("add.u32 %0, %0, %1;" : "+r"(x) : "r"(y));
  else if (K == LOP3) x = (x & y) ^ a;
  else if (K == SHF) x = __funnelshift_r(x, y, 7);
  else if (K == PRMT) x = __byte_perm(x, y, 0x6521);
  else if (K == VADD2) x = __vadd2(x, y);
  else if (K == VMNMX2) x = __vmaxs2(x, y);
  else if (K == VADDMNMX2) x = __viaddmax_s16x2(x, y, a);
  else if (K == FFMA) x = __float_as_uint(fmaf(__uint_as_float(x), __uint_as_float(a) * 0.0f + 0.5f + __uint_as_float(a), __uint_as_float(y)));
  else if (K == FADD) x = __float_as_uint(__fadd_rn(__uint_as_float(x), __uint_as_float(y)));
  else if (K == FMNMX) x = __float_as_uint(fminf(__uint_as_float(x), __uint_as_float(y)));
  else if (K == I2FP) x = __float_as_uint((float)(int)(x ^ y));
  else if (K == IDP2A) asm volatile("dp2a.lo.s32.s32 %0, %1, %2, %0;" : "+r"(x) : "r"(y), "r"(a));
  else if (K == IMNMX) x = (uint32_t)max((int)x, (int)y);
  else if (K == FMUL) x = __float_as_uint(__fmul_rn(__uint_as_float(x), __uint_as_float(y)));
  else if (K == HFMA2) asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(x) : "r"(a), "r"(y));
  else if (K == LDS64) { uint2 v = *reinterpret_cast<const uint2*>(sm + ((x & 0x3f) * 2)); x = v.x ^ v.y ^ y; }
}
template <int K0, int K1>
__global__ void __launch_bounds__(512) bench(uint32_t* out, uint32_t a, long long* cyc, unsigned long long* ns) {
  __shared__ uint32_t sm[256];
  if (threadIdx.x < 256) sm[threadIdx.x] = threadIdx.x * 3;
  uint32_t x[CH];
#pragma unroll
  for (int i = 0; i < CH; i++) x[i] = __float_as_uint(1.0f + (threadIdx.x & 31) * 0.125f + i);
  __syncthreads();
  unsigned long long n0, n1;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(n0));
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITER; it++) {
#pragma unroll
    for (int u = 0; u < UNR; u++)
#pragma unroll
      for (int i = 0; i < CH; i++) { if (i & 1) op<K1>(x[i], x[(i + 2) % CH], a, sm); else op<K0>(x[i], x[(i + 2) % CH], a, sm); }
  }
  __syncthreads();
  const long long t1 = clock64();
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(n1));
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < CH; i++) s ^= x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) { cyc[blockIdx.x] = t1 - t0; ns[blockIdx.x] = n1 - n0; }
}
static uint32_t* out; static long long* cyc; static unsigned long long* ns; static int blocks;
static double g_mhz = 0;
template <int K0, int K1> double run() {
  bench<K0, K1><<<blocks, 512>>>(out, 3, cyc, ns);
  bench<K0, K1><<<blocks, 512>>>(out, 3, cyc, ns);
  cudaDeviceSynchronize();
  static long long h[4096]; static unsigned long long hn[4096];
  cudaMemcpy(h, cyc, blocks * 8, cudaMemcpyDeviceToHost);
  cudaMemcpy(hn, ns, blocks * 8, cudaMemcpyDeviceToHost);
  double avg = 0, avgn = 0; for (int i = 0; i < blocks; i++) { avg += h[i]; avgn += hn[i]; } avg /= blocks; avgn /= blocks;
  g_mhz = avg / avgn * 1000.0;
  return 4.0 * ITER * CH * UNR / avg;
}
template <int K> void row() {
  const double alone = run<K, K>(); const double mhz = g_mhz;
  printf("%-16s alone %.3f   +IMAD %.3f  +IADD3 %.3f  +FFMA %.3f  +PRMT %.3f  +VIADD2 %.3f   (clock64 = %.0f MHz)\n", names[K], alone, run<K, IMAD>(), run<K, IADD3>(), run<K, FFMA>(), run<K, PRMT>(), run<K, VADD2>(), mhz);
}
int main() {
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  blocks = sms;  // one 512-thread block per SM: 4 warps per SMSP, whatever the register count
  cudaMalloc(&out, blocks * 512 * 4); cudaMalloc(&cyc, blocks * 8); cudaMalloc(&ns, blocks * 8);
  run<IMAD, IMAD>(); run<FFMA, FFMA>();  // warm up the clocks
  printf("warp-instructions per clock64 tick per SMSP, 4 resident warps per SMSP, 8 independent chains per thread\n");
  row<IMAD>(); row<IADD3>(); row<LOP3>(); row<SHF>(); row<PRMT>(); row<VADD2>(); row<VMNMX2>(); row<VADDMNMX2>(); row<FFMA>(); row<FADD>(); row<FMNMX>(); row<I2FP>();
  row<IDP2A>(); row<IMNMX>(); row<FMUL>(); row<HFMA2>(); row<LDS64>();
  return 0;
}
