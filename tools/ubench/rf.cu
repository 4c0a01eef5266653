// rf.cu -- is the register file's read bandwidth what caps the issue rate of IDP.2A + ALU mixes?  Every test runs 8 independent
// chains per thread, 4 warps per SMSP, loops of 64 instructions (L0-resident); operands are distinct registers unless stated.
//   iadd3_3r   d = a + b + c, three registers         iadd_2r  d = a + b         iadd_1r  d = a + imm
//   idp_3r     IDP.2A, three registers                 idp_2r   IDP.2A with the coefficient word in a uniform register
//   A+B        the two kinds alternate 1:1, TOTAL instructions per clock per SMSP
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o rf rf.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
constexpr int ITER = 2048;
enum { IADD3_3R, IADD_2R, IADD_1R, IDP_3R, IDP_2R, LOP3_3R, VMNMX_2R, SHF_2R, PRMT_2R, NK };
const char* nm[NK] = {"iadd3_3r", "iadd_2r", "iadd_1r", "idp_3r", "idp_2r", "lop3_3r", "vimnmx2_2r", "shf_2r", "prmt_2r"};
template <int K>
__device__ __forceinline__ void op(uint32_t& d, uint32_t a, uint32_t b, uint32_t u) {
  if (K == IADD3_3R) asm volatile("{.reg .u32 t; add.u32 t, %1, %2; add.u32 %0, t, %0;}" : "+r"(d) : "r"(a), "r"(b));
  else if (K == IADD_2R) asm volatile("add.u32 %0, %0, %1;" : "+r"(d) : "r"(a));
  else if (K == IADD_1R) asm volatile("add.u32 %0, %0, 12345;" : "+r"(d));
  else if (K == IDP_3R) asm volatile("dp2a.lo.u32.s32 %0, %1, %2, %0;" : "+r"(d) : "r"(a), "r"(b));
  else if (K == IDP_2R) asm volatile("dp2a.lo.u32.s32 %0, %1, %2, %0;" : "+r"(d) : "r"(a), "r"(u));
  else if (K == LOP3_3R) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(d) : "r"(a), "r"(b));
  else if (K == VMNMX_2R) d = __vmaxu2(d, a);
  else if (K == SHF_2R) d = __funnelshift_r(d, a, 7);
  else if (K == PRMT_2R) d = __byte_perm(d, a, 0x6521);
}
template <int K0, int K1>
__global__ void __launch_bounds__(512) bench(uint32_t* out, uint32_t seed, long long* cyc) {
  uint32_t d[8], a[8], b[8];
#pragma unroll
  for (int i = 0; i < 8; i++) { d[i] = seed + i; a[i] = seed * (i + 3) + threadIdx.x; b[i] = seed * (i + 17) ^ threadIdx.x; }
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITER; it++) {
#pragma unroll
    for (int u = 0; u < 4; u++)
#pragma unroll
      for (int i = 0; i < 8; i++) { op<K0>(d[i], a[(i + u) & 7], b[(i + 2 * u + 1) & 7], seed); op<K1>(d[(i + 4) & 7], b[(i + u) & 7], a[(i + 3 * u + 2) & 7], seed); }
  }
  const long long t1 = clock64();
  uint32_t r = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) r ^= d[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
static uint32_t* out; static long long* cyc; static int sms;
template <int K0, int K1> double run() {
  bench<K0, K1><<<sms, 512>>>(out, 3, cyc); bench<K0, K1><<<sms, 512>>>(out, 3, cyc);
  cudaDeviceSynchronize();
  static long long h[1024]; cudaMemcpy(h, cyc, sms * 8, cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < sms; i++) avg += h[i]; avg /= sms;
  return 4.0 * ITER * 64 / avg;
}
template <int K> void row() {
  printf("%-11s alone %.3f  +idp_3r %.3f  +idp_2r %.3f  +iadd3_3r %.3f  +iadd_2r %.3f  +iadd_1r %.3f  +lop3_3r %.3f  +vimnmx2 %.3f  +shf %.3f  +prmt %.3f\n", nm[K], run<K, K>(), run<K, IDP_3R>(),
         run<K, IDP_2R>(), run<K, IADD3_3R>(), run<K, IADD_2R>(), run<K, IADD_1R>(), run<K, LOP3_3R>(), run<K, VMNMX_2R>(), run<K, SHF_2R>(), run<K, PRMT_2R>());
}
int main() {
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaMalloc(&out, sms * 512 * 4); cudaMalloc(&cyc, sms * 8);
  run<IDP_3R, IDP_3R>();
  printf("warp-instructions per clock per SMSP (iadd3_3r counts as ONE instruction only if the two adds fuse: check SASS)\n");
  row<IDP_3R>(); row<IDP_2R>(); row<IADD3_3R>(); row<IADD_2R>(); row<IADD_1R>(); row<LOP3_3R>(); row<VMNMX_2R>(); row<SHF_2R>(); row<PRMT_2R>();
  return 0;
}
