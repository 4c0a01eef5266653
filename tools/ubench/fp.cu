// fp.cu -- issue rate of the fp32 instructions considered for the ALF FIR (sm_100a): FFMA, FFMA2 (fma.rn.f32x2), FADD.RM,
// I2F/F2I and their 1:1 mixes with integer ALU work.  Prints source operations per clock per SMSP (1.0 = full issue rate
// for single-instruction operations).
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp fp.cu ; run on the GPU box.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
constexpr int ITER = 512, CH = 8, UNR = 8;
enum { FFMA, FFMA2, FFMA2_LOP, FFMA_LOP, FFMA_IMAD, I2F, F2I, FADDRM, IMAD, IMAD_LOP, PRMT_FADD2, DP2A, DP2A_LOP, DP2A_IMAD, NK };
const char* names[NK] = {"FFMA", "FFMA2", "FFMA2+LOP3 1:1", "FFMA+LOP3 1:1", "FFMA+IMAD 1:1", "I2F.S32", "F2I.FLOOR", "FADD.RM", "IMAD", "IMAD+LOP3 1:1", "2xPRMT+FADD2", "IDP.2A", "IDP.2A+LOP3 1:1", "IDP.2A+IMAD 1:1"};
template <int K>
__global__ void __launch_bounds__(512) bench(uint32_t* out, float a, float b, uint32_t ia, long long* cyc) {
  unsigned long long x[CH];
  uint32_t z[CH];
#pragma unroll
  for (int i = 0; i < CH; i++) { x[i] = ((unsigned long long)__float_as_uint(1.0f + threadIdx.x + i) << 32) | __float_as_uint(0.5f + i); z[i] = threadIdx.x + i; }
  unsigned long long aa = ((unsigned long long)__float_as_uint(a) << 32) | __float_as_uint(a), bb = ((unsigned long long)__float_as_uint(b) << 32) | __float_as_uint(b);
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITER; it++) {
#pragma unroll
    for (int u = 0; u < UNR; u++)
#pragma unroll
      for (int i = 0; i < CH; i++) {
        if (K == FFMA) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+r"(z[i]) : "r"(__float_as_uint(a)), "r"(__float_as_uint(b)));
        else if (K == FFMA2) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[i]) : "l"(aa), "l"(bb));
        else if (K == FFMA2_LOP) { if (i & 1) z[i] = (z[i] & ia) ^ 0x5410; else asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[i]) : "l"(aa), "l"(bb)); }
        else if (K == FFMA_LOP) { if (i & 1) z[i] = (z[i] & ia) ^ 0x5410; else asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+r"(z[i]) : "r"(__float_as_uint(a)), "r"(__float_as_uint(b))); }
        else if (K == FFMA_IMAD) { if (i & 1) z[i] = z[i] * ia + 0x5410; else asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+r"(z[i]) : "r"(__float_as_uint(a)), "r"(__float_as_uint(b))); }
        else if (K == I2F) z[i] = __float_as_uint((float)(int)z[i]);
        else if (K == F2I) z[i] = (uint32_t)__float2int_rd(__uint_as_float(z[i]));
        else if (K == FADDRM) z[i] = __float_as_uint(__fadd_rd(__uint_as_float(z[i]), a));
        else if (K == IMAD) z[i] = z[i] * ia + 0x5410;
        else if (K == IMAD_LOP) { if (i & 1) z[i] = (z[i] & ia) ^ 0x5410; else z[i] = z[i] * ia + 0x5410; }
        else if (K == DP2A) asm volatile("dp2a.lo.s32.s32 %0, %0, %1, %2;" : "+r"(z[i]) : "r"(ia), "r"(z[i]));
        else if (K == DP2A_LOP) { if (i & 1) z[i] = (z[i] & ia) ^ 0x5410; else asm volatile("dp2a.lo.s32.s32 %0, %0, %1, %2;" : "+r"(z[i]) : "r"(ia), "r"(z[i])); }
        else if (K == DP2A_IMAD) { if (i & 1) z[i] = z[i] * ia + 0x5410; else asm volatile("dp2a.lo.s32.s32 %0, %0, %1, %2;" : "+r"(z[i]) : "r"(ia), "r"(z[i])); }
        else if (K == PRMT_FADD2) {
          // two magic-float builds + one packed subtract: int16 pair -> two fp32
          uint32_t lo = __byte_perm(z[i], 0x4B000000u, 0x7410), hi = __byte_perm(z[i], 0x4B000000u, 0x7432);
          unsigned long long p = ((unsigned long long)hi << 32) | lo;
          asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(p) : "l"(p), "l"(bb));
          z[i] = (uint32_t)p ^ (uint32_t)(p >> 32);
        }
      }
  }
  __syncthreads();
  const long long t1 = clock64();
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < CH; i++) s ^= z[i] ^ (uint32_t)x[i] ^ (uint32_t)(x[i] >> 32);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
static uint32_t* out; static long long* cyc; static int blocks;
template <int K> void run() {
  bench<K><<<blocks, 512>>>(out, 1.0001f, 0.5f, 3, cyc);
  bench<K><<<blocks, 512>>>(out, 1.0001f, 0.5f, 3, cyc);
  cudaDeviceSynchronize();
  static long long h[4096];
  cudaMemcpy(h, cyc, blocks * 8, cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < blocks; i++) avg += h[i]; avg /= blocks;
  printf("%-18s %.3f source-ops per clock per SMSP\n", names[K], 16.0 * ITER * CH * UNR / avg);
}
int main() {
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  blocks = sms * 4;
  cudaMalloc(&out, blocks * 512 * 4); cudaMalloc(&cyc, blocks * 8);
  run<FFMA>(); run<FFMA2>(); run<FFMA2_LOP>(); run<FFMA_LOP>(); run<FFMA_IMAD>(); run<I2F>(); run<F2I>(); run<FADDRM>(); run<IMAD>(); run<IMAD_LOP>(); run<PRMT_FADD2>();
  run<DP2A>(); run<DP2A_LOP>(); run<DP2A_IMAD>();
  return 0;
}
