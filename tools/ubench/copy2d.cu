// copy2d.cu -- which data-movement skeleton moves a 4:2:0 int16 picture batch through an SM fastest? (sm_100a)
// Same tile shapes as the deblocking kernel (128x32 luma + 2 x 64x16 chroma per CTA); no arithmetic.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int W = 3840, H = 2160, NS = 17;
constexpr int PY = 3840, PC = 1920;
constexpr size_t PLANE_Y = (size_t)PY * H, PLANE_C = (size_t)PC * (H / 2), BUF = PLANE_Y + 2 * PLANE_C;

struct P { const int16_t* src; int16_t* dst; };

__global__ void __launch_bounds__(256) v0_flat(const uint4* __restrict__ s, uint4* __restrict__ d, size_t n) {
  size_t i = (size_t)blockIdx.x * 256 * 4 + threadIdx.x;
  uint4 v[4];
#pragma unroll
  for (int k = 0; k < 4; k++) if (i + k * 256 < n) v[k] = __ldg(s + i + k * 256);
#pragma unroll
  for (int k = 0; k < 4; k++) if (i + k * 256 < n) d[i + k * 256] = v[k];
}

// aligned tile, registers only
__global__ void __launch_bounds__(128) v1_tile_regs(const int16_t* __restrict__ src, int16_t* __restrict__ dst) {
  const int16_t* s = src + (size_t)blockIdx.z * BUF; int16_t* d = dst + (size_t)blockIdx.z * BUF;
  const int tid = threadIdx.x, x0 = blockIdx.x * 128, y0 = blockIdx.y * 32;
  uint4 ly[4], lc[2];
  const int k = tid & 15, r0 = tid >> 4;
#pragma unroll
  for (int i = 0; i < 4; i++) { const int y = y0 + r0 + 8 * i; if (y < H) ly[i] = __ldg((const uint4*)(s + (size_t)y * PY + x0 + 8 * k)); }
  const int kc = tid & 7, rc = (tid >> 3) & 7, pl = 0;
#pragma unroll
  for (int i = 0; i < 2; i++) {
    const int plane = (tid >> 6), y = blockIdx.y * 16 + rc + 8 * i;
    if (y < H / 2) lc[i] = __ldg((const uint4*)(s + PLANE_Y + plane * PLANE_C + (size_t)y * PC + blockIdx.x * 64 + 8 * kc));
  }
#pragma unroll
  for (int i = 0; i < 4; i++) { const int y = y0 + r0 + 8 * i; if (y < H) *(uint4*)(d + (size_t)y * PY + x0 + 8 * k) = ly[i]; }
#pragma unroll
  for (int i = 0; i < 2; i++) {
    const int plane = (tid >> 6), y = blockIdx.y * 16 + rc + 8 * i;
    if (y < H / 2) *(uint4*)(d + PLANE_Y + plane * PLANE_C + (size_t)y * PC + blockIdx.x * 64 + 8 * kc) = lc[i];
  }
  (void)pl;
}

// aligned tile through shared memory
__global__ void __launch_bounds__(128) v2_tile_smem(const int16_t* __restrict__ src, int16_t* __restrict__ dst) {
  __shared__ __align__(16) int16_t sy[32 * 136];
  __shared__ __align__(16) int16_t sc[2][16 * 72];
  const int16_t* s = src + (size_t)blockIdx.z * BUF; int16_t* d = dst + (size_t)blockIdx.z * BUF;
  const int tid = threadIdx.x, x0 = blockIdx.x * 128, y0 = blockIdx.y * 32;
  uint4 ly[4], lc[2];
  const int k = tid & 15, r0 = tid >> 4;
  const int kc = tid & 7, rc = (tid >> 3) & 7, plane = tid >> 6;
#pragma unroll
  for (int i = 0; i < 4; i++) { const int y = y0 + r0 + 8 * i; ly[i] = make_uint4(0, 0, 0, 0); if (y < H) ly[i] = __ldg((const uint4*)(s + (size_t)y * PY + x0 + 8 * k)); }
#pragma unroll
  for (int i = 0; i < 2; i++) {
    const int y = blockIdx.y * 16 + rc + 8 * i; lc[i] = make_uint4(0, 0, 0, 0);
    if (y < H / 2) lc[i] = __ldg((const uint4*)(s + PLANE_Y + plane * PLANE_C + (size_t)y * PC + blockIdx.x * 64 + 8 * kc));
  }
#pragma unroll
  for (int i = 0; i < 4; i++) *(uint4*)&sy[(r0 + 8 * i) * 136 + 8 * k] = ly[i];
#pragma unroll
  for (int i = 0; i < 2; i++) *(uint4*)&sc[plane][(rc + 8 * i) * 72 + 8 * kc] = lc[i];
  __syncthreads();
  // rotate ownership so that the write-back really reads other threads' data
  const int t2 = (tid + 37) & 127;
  const int k2 = t2 & 15, r2 = t2 >> 4, kc2 = t2 & 7, rc2 = (t2 >> 3) & 7, plane2 = t2 >> 6;
#pragma unroll
  for (int i = 0; i < 4; i++) { const int y = y0 + r2 + 8 * i; if (y < H) *(uint4*)(d + (size_t)y * PY + x0 + 8 * k2) = *(uint4*)&sy[(r2 + 8 * i) * 136 + 8 * k2]; }
#pragma unroll
  for (int i = 0; i < 2; i++) {
    const int y = blockIdx.y * 16 + rc2 + 8 * i;
    if (y < H / 2) *(uint4*)(d + PLANE_Y + plane2 * PLANE_C + (size_t)y * PC + blockIdx.x * 64 + 8 * kc2) = *(uint4*)&sc[plane2][(rc2 + 8 * i) * 72 + 8 * kc2];
  }
}

// tile shifted by (-4,-4) luma / (-4,-2) chroma, 8-byte accesses, through shared memory (the deblocking kernel's skeleton)
template <int XS>
__global__ void __launch_bounds__(128) v3_shift_smem(const int16_t* __restrict__ src, int16_t* __restrict__ dst) {
  __shared__ __align__(16) int16_t sy[32 * 136];
  __shared__ __align__(16) int16_t sc[2][16 * 72];
  const int16_t* s = src + (size_t)blockIdx.z * BUF; int16_t* d = dst + (size_t)blockIdx.z * BUF;
  const int tid = threadIdx.x, x0 = blockIdx.x * 128 - XS, y0 = blockIdx.y * 32 - 4, cx0 = blockIdx.x * 64 - XS, cy0 = blockIdx.y * 16 - 2;
  uint2 ly[8], lc[4];
  const int k = tid & 31, r0 = tid >> 5, kc = tid & 15, rc0 = tid >> 4;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const int x = x0 + 4 * k, y = y0 + r0 + 4 * i; ly[i] = make_uint2(0, 0);
    if (x >= 0 && x < W && y >= 0 && y < H) ly[i] = __ldg((const uint2*)(s + (size_t)y * PY + x));
  }
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int pl = i >> 1, x = cx0 + 4 * kc, y = cy0 + rc0 + 8 * (i & 1); lc[i] = make_uint2(0, 0);
    if (x >= 0 && x < W / 2 && y >= 0 && y < H / 2) lc[i] = __ldg((const uint2*)(s + PLANE_Y + pl * PLANE_C + (size_t)y * PC + x));
  }
#pragma unroll
  for (int i = 0; i < 8; i++) *(uint2*)&sy[(r0 + 4 * i) * 136 + 4 * k] = ly[i];
#pragma unroll
  for (int i = 0; i < 4; i++) *(uint2*)&sc[i >> 1][(rc0 + 8 * (i & 1)) * 72 + 4 * kc] = lc[i];
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const int x = x0 + 4 * k, y = y0 + r0 + 4 * i;
    if (x >= 0 && x < W && y >= 0 && y < H) *(uint2*)(d + (size_t)y * PY + x) = *(uint2*)&sy[(r0 + 4 * i) * 136 + 4 * k];
  }
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int pl = i >> 1, x = cx0 + 4 * kc, y = cy0 + rc0 + 8 * (i & 1);
    if (x >= 0 && x < W / 2 && y >= 0 && y < H / 2) *(uint2*)(d + PLANE_Y + pl * PLANE_C + (size_t)y * PC + x) = *(uint2*)&sc[pl][(rc0 + 8 * (i & 1)) * 72 + 4 * kc];
  }
}

// aligned 16-byte tile + read-only 8-byte halo columns left and right, rows shifted by -4 / -2, through shared memory
__global__ void __launch_bounds__(128) v4_halo_smem(const int16_t* __restrict__ src, int16_t* __restrict__ dst) {
  __shared__ __align__(16) int16_t sy[32 * 152];     // [8 halo-left slot][128][8 halo-right slot] -> data at +8
  __shared__ __align__(16) int16_t sc[2][16 * 88];
  const int16_t* s = src + (size_t)blockIdx.z * BUF; int16_t* d = dst + (size_t)blockIdx.z * BUF;
  const int tid = threadIdx.x, x0 = blockIdx.x * 128, y0 = blockIdx.y * 32 - 4, cx0 = blockIdx.x * 64, cy0 = blockIdx.y * 16 - 2;
  uint4 ly[4], lc[2];
  uint2 hy = make_uint2(0, 0), hc = make_uint2(0, 0);
  const int k = tid & 15, r0 = tid >> 4, kc = tid & 7, rc = (tid >> 3) & 7, plane = tid >> 6;
#pragma unroll
  for (int i = 0; i < 4; i++) { const int y = y0 + r0 + 8 * i; ly[i] = make_uint4(0, 0, 0, 0); if (y >= 0 && y < H) ly[i] = __ldg((const uint4*)(s + (size_t)y * PY + x0 + 8 * k)); }
#pragma unroll
  for (int i = 0; i < 2; i++) {
    const int y = cy0 + rc + 8 * i; lc[i] = make_uint4(0, 0, 0, 0);
    if (y >= 0 && y < H / 2) lc[i] = __ldg((const uint4*)(s + PLANE_Y + plane * PLANE_C + (size_t)y * PC + cx0 + 8 * kc));
  }
  {  // halo: luma 32 rows x 2 sides = 64 threads, chroma 2 planes x 16 rows x 2 sides = 64 threads
    if (tid < 64) {
      const int r = tid >> 1, side = tid & 1, y = y0 + r, x = side ? x0 + 128 : x0 - 4;
      if (y >= 0 && y < H && x >= 0 && x < W) hy = __ldg((const uint2*)(s + (size_t)y * PY + x));
    } else {
      const int t = tid - 64, pl = t >> 5, r = (t >> 1) & 15, side = t & 1, y = cy0 + r, x = side ? cx0 + 64 : cx0 - 4;
      if (y >= 0 && y < H / 2 && x >= 0 && x < W / 2) hc = __ldg((const uint2*)(s + PLANE_Y + pl * PLANE_C + (size_t)y * PC + x));
    }
  }
#pragma unroll
  for (int i = 0; i < 4; i++) *(uint4*)&sy[(r0 + 8 * i) * 152 + 8 + 8 * k] = ly[i];
#pragma unroll
  for (int i = 0; i < 2; i++) *(uint4*)&sc[plane][(rc + 8 * i) * 88 + 8 + 8 * kc] = lc[i];
  if (tid < 64) { const int r = tid >> 1, side = tid & 1; *(uint2*)&sy[r * 152 + (side ? 136 : 4)] = hy; }
  else { const int t = tid - 64, pl = t >> 5, r = (t >> 1) & 15, side = t & 1; *(uint2*)&sc[pl][r * 88 + (side ? 72 : 4)] = hc; }
  __syncthreads();
  const int t2 = (tid + 37) & 127;
  const int k2 = t2 & 15, r2 = t2 >> 4, kc2 = t2 & 7, rc2 = (t2 >> 3) & 7, plane2 = t2 >> 6;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int y = y0 + r2 + 8 * i;
    uint4 v = *(uint4*)&sy[(r2 + 8 * i) * 152 + 8 + 8 * k2];
    v.x ^= sy[(r2 + 8 * i) * 152 + 4 + k2] & 0;  // touch the halo
    if (y >= 0 && y < H) *(uint4*)(d + (size_t)y * PY + x0 + 8 * k2) = v;
  }
#pragma unroll
  for (int i = 0; i < 2; i++) {
    const int y = cy0 + rc2 + 8 * i;
    if (y >= 0 && y < H / 2) *(uint4*)(d + PLANE_Y + plane2 * PLANE_C + (size_t)y * PC + cx0 + 8 * kc2) = *(uint4*)&sc[plane2][(rc2 + 8 * i) * 88 + 8 + 8 * kc2];
  }
}

template <typename F>
void timeit(const char* name, F launch, double bytes) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int i = 0; i < 3; i++) launch();
  cudaEventRecord(a);
  const int N = 20;
  for (int i = 0; i < N; i++) launch();
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  printf("%-40s %8.3f ms  %8.1f GB/s  %s\n", name, ms / N, bytes / (ms / N * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  int16_t *src, *dst;
  cudaMalloc(&src, NS * BUF * 2); cudaMalloc(&dst, NS * BUF * 2);
  cudaMemset(src, 1, NS * BUF * 2); cudaMemset(dst, 0, NS * BUF * 2);
  const double bytes = 2.0 * NS * BUF * 2;
  const size_t n16 = NS * BUF * 2 / 16;
  timeit("cudaMemcpyAsync D2D", [&] { cudaMemcpyAsync(dst, src, NS * BUF * 2, cudaMemcpyDeviceToDevice); }, bytes);
  timeit("v0 flat 16B x4 per thread", [&] { v0_flat<<<(unsigned)((n16 + 1023) / 1024), 256>>>((const uint4*)src, (uint4*)dst, n16); }, bytes);
  dim3 g(30, 68, NS), gs(31, 68, NS);
  timeit("v1 aligned tile, registers", [&] { v1_tile_regs<<<g, 128>>>(src, dst); }, bytes);
  timeit("v2 aligned tile, smem", [&] { v2_tile_smem<<<g, 128>>>(src, dst); }, bytes);
  timeit("v3 shifted(-4,-4) 8B, smem [current]", [&] { v3_shift_smem<4><<<gs, 128>>>(src, dst); }, bytes);
  timeit("v3b shifted(0,-4) 8B, smem", [&] { v3_shift_smem<0><<<gs, 128>>>(src, dst); }, bytes);
  timeit("v4 aligned 16B + 8B halos, rows -4, smem", [&] { v4_halo_smem<<<dim3(30, 68, NS), 128>>>(src, dst); }, bytes);
  return 0;
}
