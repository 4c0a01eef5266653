// mix.cu -- what issue rate can a stream like the ALF luma filter reach?  IDP.2A with three distinct register operands
// (window word, coefficient word, accumulator) and ALU work (IADD3 / VIMNMX.U16x2 / SHF on other registers), arranged
//   idp      : IDP only                                   alu : ALU only
//   fine     : one ALU instruction after every IDP        coarse: NB IDP, then NB ALU (same warp)
//   phased   : as coarse, but odd warps start with the ALU block (two warps of an SMSP are always in different blocks)
// Prints warp-instructions per clock per SMSP for 4, 5 and 8 warps per SMSP.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o mix mix.cu ; run on the GPU box.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
constexpr int ITER = 256;
__device__ __forceinline__ int dp(uint32_t a, uint32_t b, int c, int hi) {
  int d;
  if (hi) asm volatile("dp2a.hi.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  else asm volatile("dp2a.lo.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
// one "unit" of IDP work: 32 IDPs over 8 accumulators, 12 window words, 8 coefficient words
__device__ __forceinline__ void idp_unit(int (&acc)[8], const uint32_t (&w)[12], const uint32_t (&c)[8]) {
#pragma unroll
  for (int t = 0; t < 4; t++)
#pragma unroll
    for (int k = 0; k < 8; k++) acc[k] = dp(w[(k + 3 * t) % 12], c[(k + t) & 7], acc[k], (k + t) & 1);
}
// one unit of ALU work: 32 instructions (IADD3, IADD, VIMNMX.U16x2, IADD) x 8 chains, like the Laplacian
__device__ __forceinline__ void alu_unit(uint32_t (&s)[8], const uint32_t (&m)[8], uint32_t k2) {
#pragma unroll
  for (int k = 0; k < 8; k++) {
    uint32_t u;
    asm volatile("{.reg .u32 t; add.u32 t, %1, %2; sub.u32 %0, t, %3;}" : "=r"(u) : "r"(m[k]), "r"(m[(k + 1) & 7]), "r"(m[(k + 3) & 7]));
    uint32_t n; asm volatile("sub.u32 %0, %1, %2;" : "=r"(n) : "r"(k2), "r"(u));
    const uint32_t v = __vmaxu2(u, n);
    asm volatile("add.u32 %0, %0, %1;" : "+r"(s[k]) : "r"(v));
  }
}
template <int MODE, int NB>
__global__ void bench(uint32_t* out, uint32_t seed, long long* cyc) {
  int acc[8]; uint32_t w[12], c[8], s[8], m[8];
#pragma unroll
  for (int i = 0; i < 8; i++) { acc[i] = i; c[i] = seed * (i + 3) + threadIdx.x; s[i] = i; m[i] = seed * (i + 11) ^ threadIdx.x; }
#pragma unroll
  for (int i = 0; i < 12; i++) w[i] = seed * (i + 7) + threadIdx.x * 5;
  const bool odd = (threadIdx.x >> 5) & 1;
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITER; it++) {
    if (MODE == 0) {
#pragma unroll
      for (int b = 0; b < 2 * NB; b++) idp_unit(acc, w, c);
    } else if (MODE == 1) {
#pragma unroll
      for (int b = 0; b < 2 * NB; b++) alu_unit(s, m, seed);
    } else if (MODE == 2) {  // fine: interleave instruction by instruction (8 chains each)
#pragma unroll
      for (int b = 0; b < NB; b++) {
#pragma unroll
        for (int t = 0; t < 4; t++)
#pragma unroll
          for (int k = 0; k < 8; k++) {
            acc[k] = dp(w[(k + 3 * t) % 12], c[(k + t) & 7], acc[k], (k + t) & 1);
            if (t == 0) { asm volatile("{.reg .u32 t; add.u32 t, %1, %2; sub.u32 %0, t, %3;}" : "=r"(m[k]) : "r"(s[k]), "r"(m[(k + 1) & 7]), "r"(m[(k + 3) & 7])); }
            else if (t == 1) asm volatile("sub.u32 %0, %1, %2;" : "=r"(s[k]) : "r"(seed), "r"(m[k]));
            else if (t == 2) s[k] = __vmaxu2(s[k], m[k]);
            else asm volatile("add.u32 %0, %0, %1;" : "+r"(m[k]) : "r"(s[k]));
          }
      }
    } else if (MODE == 3) {  // coarse
#pragma unroll
      for (int b = 0; b < NB; b++) idp_unit(acc, w, c);
#pragma unroll
      for (int b = 0; b < NB; b++) alu_unit(s, m, seed);
    } else {  // phased
      if (odd) {
#pragma unroll
        for (int b = 0; b < NB; b++) alu_unit(s, m, seed);
#pragma unroll
        for (int b = 0; b < NB; b++) idp_unit(acc, w, c);
      } else {
#pragma unroll
        for (int b = 0; b < NB; b++) idp_unit(acc, w, c);
#pragma unroll
        for (int b = 0; b < NB; b++) alu_unit(s, m, seed);
      }
    }
  }
  const long long t1 = clock64();
  uint32_t r = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) r ^= acc[i] ^ s[i] ^ m[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
static uint32_t* out; static long long* cyc; static int sms;
template <int MODE, int NB> double run(int threads) {
  bench<MODE, NB><<<sms, threads>>>(out, 3, cyc);
  bench<MODE, NB><<<sms, threads>>>(out, 3, cyc);
  cudaDeviceSynchronize();
  static long long h[1024];
  cudaMemcpy(h, cyc, sms * 8, cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < sms; i++) avg += h[i]; avg /= sms;
  return (threads / 128.0) * ITER * 2.0 * NB * 32 / avg;  // warps per SMSP x instructions per warp / cycles
}
template <int NB> void rows() {
  const char* nm[5] = {"idp", "alu", "fine", "coarse", "phased"};
  for (int th : {512, 640, 1024}) {
    printf("block %2d units, %d warps/SMSP:", NB, th / 128);
    printf("  %s %.3f", nm[0], run<0, NB>(th)); printf("  %s %.3f", nm[1], run<1, NB>(th)); printf("  %s %.3f", nm[2], run<2, NB>(th));
    printf("  %s %.3f", nm[3], run<3, NB>(th)); printf("  %s %.3f\n", nm[4], run<4, NB>(th));
  }
}
int main() {
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaMalloc(&out, sms * 1024 * 4); cudaMalloc(&cyc, sms * 8);
  run<0, 1>(512);
  rows<1>(); rows<4>(); rows<16>(); rows<32>(); rows<64>();
  return 0;
}
