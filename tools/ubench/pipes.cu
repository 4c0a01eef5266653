// pipes.cu -- instruction-throughput microbenchmark for the integer instructions the filter kernels lean on (sm_100a).
// Prints warp-instructions per clock per SM sub-partition (SMSP) for each instruction kind; 1.0 = full issue rate.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o pipes pipes.cu ; run on the GPU box.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int ITER = 2048, CH = 8;

template <int KIND>
__device__ __forceinline__ void op(uint32_t& x, uint32_t a, uint32_t b) {
  if (KIND == 0) x = x * a + b;                                            // IMAD
  else if (KIND == 1) asm volatile("add.u32 %0, %0, %1;" : "+r"(x) : "r"(a));  // IADD3
  else if (KIND == 2) x = __vadd2(x, a);                                   // VIADD.16x2 ?
  else if (KIND == 3) x = __vmaxs2(x, a);                                  // VIMNMX.S16x2
  else if (KIND == 4) x = __viaddmin_s16x2_relu(x, a, b);                  // VIADDMNMX.S16x2.RELU
  else if (KIND == 5) x = __byte_perm(x, a, 0x5410 ^ (b & 0x1111));        // PRMT (reg selector)
  else if (KIND == 6) asm volatile("dp2a.lo.s32.s32 %0, %0, %1, %2;" : "+r"(x) : "r"(a), "r"(b));   // IDP.2A
  else if (KIND == 7) asm volatile("dp4a.s32.s32 %0, %0, %1, %2;" : "+r"(x) : "r"(a), "r"(b));      // IDP.4A
  else if (KIND == 8) x = (x & a) ^ b;                                     // LOP3
  else if (KIND == 9) x = __funnelshift_r(x, a, 16);                       // SHF
  else if (KIND == 10) x = __vabsdiffs2(x, a);                             // emulated?
  else if (KIND == 11) x = __vsub2(x, a);
  else if (KIND == 12) asm volatile("dp2a.lo.u32.s32 %0, %0, %1, %2;" : "+r"(x) : "r"(a), "r"(b));
  else if (KIND == 13) x = max(min((int)x, (int)a), (int)b);               // VIMNMX3 ?
  else if (KIND == 14) x = (uint32_t)abs((int)x - (int)a);                 // IABS + IADD
  else if (KIND == 15) x = __vabs2(x);
  else if (KIND == 16) x = __vneg2(x);
  else if (KIND == 17) x = __vcmpgts2(x, a);
  else if (KIND == 18) x = __vhaddu2(x, a);
}

// MIX: two kinds interleaved 1:1 on independent chains
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
    for (int i = 0; i < CH; i++) {
      if (i & 1) op<K1>(x[i], a, b); else op<K0>(x[i], a, b);
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < CH; i++) s ^= x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int K0, int K1>
void run(const char* name) {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int blocks = sms * 4;  // 4 x 512 threads = 64 warps per SM (full occupancy if registers allow)
  uint32_t* out; long long* cyc;
  cudaMalloc(&out, blocks * 512 * 4); cudaMalloc(&cyc, blocks * 8);
  bench<K0, K1><<<blocks, 512>>>(out, 3, 5, cyc);
  bench<K0, K1><<<blocks, 512>>>(out, 3, 5, cyc);
  cudaDeviceSynchronize();
  long long* h = new long long[blocks];
  cudaMemcpy(h, cyc, blocks * 8, cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < blocks; i++) avg += h[i]; avg /= blocks;
  // per SMSP: 16 warps resident (64 per SM / 4), each issuing ITER*CH instructions of interest
  const double ipc = 16.0 * ITER * CH / avg;
  printf("%-40s %.3f warp-inst/clk/SMSP (cycles %.0f) %s\n", name, ipc, avg, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out); cudaFree(cyc); delete[] h;
}

int main() {
  run<0, 0>("IMAD");
  run<1, 1>("IADD");
  run<2, 2>("vadd2");
  run<11, 11>("vsub2");
  run<3, 3>("vmaxs2");
  run<4, 4>("viaddmin_s16x2_relu");
  run<5, 5>("PRMT(reg sel)");
  run<6, 6>("dp2a.lo.s32.s32");
  run<12, 12>("dp2a.lo.u32.s32");
  run<7, 7>("dp4a");
  run<8, 8>("LOP3");
  run<9, 9>("SHF funnel");
  run<10, 10>("vabsdiffs2");
  run<13, 13>("min/max clamp");
  run<14, 14>("abs(x-a)");
  run<15, 15>("vabs2");
  run<16, 16>("vneg2");
  run<17, 17>("vcmpgts2");
  run<18, 18>("vhaddu2");
  run<0, 1>("IMAD + IADD 1:1");
  run<0, 2>("IMAD + vadd2 1:1");
  run<6, 2>("dp2a + vadd2 1:1");
  run<6, 0>("dp2a + IMAD 1:1");
  run<2, 3>("vadd2 + vmaxs2 1:1");
  run<2, 5>("vadd2 + PRMT 1:1");
  run<2, 8>("vadd2 + LOP3 1:1");
  return 0;
}
