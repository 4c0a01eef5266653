// tma_copy.cu -- persistent TMA-pipelined tile copy with the deblocking kernel's tile shapes (sm_100a).
// Each CTA walks a static list of tiles; a ring of S stages in shared memory is filled by cp.async.bulk.tensor loads
// issued ahead by thread 0 and drained by cp.async.bulk.tensor stores.  Reports GB/s for several (S, CTAs/SM).
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

constexpr int W = 3840, H = 2160, NS = 17;
constexpr int PY = 3840, PC = 1920;
constexpr size_t PLANE_Y = (size_t)PY * H, PLANE_C = (size_t)PC * (H / 2), BUF = PLANE_Y + 2 * PLANE_C;
constexpr int STAGE_BYTES = 128 * 32 * 2 + 2 * 64 * 16 * 2;  // 12288

__device__ __forceinline__ uint32_t s2u(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s2u(b)), "r"(n)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s2u(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile(
      "{\n.reg .pred p;\nWAIT:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE;\nbra WAIT;\nDONE:\n}" ::"r"(s2u(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load3(void* dst, const CUtensorMap* m, uint64_t* bar, int x, int y, int z) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(s2u(dst)), "l"(m), "r"(s2u(bar)),
               "r"(x), "r"(y), "r"(z)
               : "memory");
}
__device__ __forceinline__ void tma_store3(const CUtensorMap* m, const void* src, int x, int y, int z) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(m), "r"(s2u(src)), "r"(x), "r"(y), "r"(z) : "memory");
}

struct Maps { CUtensorMap sy, scb, scr, dy, dcb, dcr; };

template <int S, int MODE>  // MODE 0: pure TMA; 1: threads read+modify+write the tile in smem (2 syncs)
__global__ void __launch_bounds__(128) tma_copy(const Maps* __restrict__ mp, int tiles_x, int tiles_y, int total) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + S * STAGE_BYTES);
  const int tid = threadIdx.x;
  if (tid == 0 && blockIdx.x == 0 && (s2u(smem) & 1023)) printf("smem base %x\n", s2u(smem));
  if (tid == 0) {
    for (int i = 0; i < S; i++) mbar_init(&full[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int per_slot = tiles_x * tiles_y;
  auto issue = [&](int t, int stage) {
    const int slot = t / per_slot, r = t % per_slot, ty = r / tiles_x, tx = r % tiles_x;
    uint8_t* base = smem + stage * STAGE_BYTES;
    mbar_expect(&full[stage], STAGE_BYTES);
    tma_load3(base, &mp->sy, &full[stage], tx * 128 - 4, ty * 32 - 4, slot);
    tma_load3(base + 8192, &mp->scb, &full[stage], tx * 64 - 4, ty * 16 - 2, slot);
    tma_load3(base + 8192 + 2048, &mp->scr, &full[stage], tx * 64 - 4, ty * 16 - 2, slot);
  };
  int n = 0;
  for (int t = blockIdx.x; t < total; t += gridDim.x) n++;
  if (tid == 0)
    for (int i = 0; i < S - 1 && i < n; i++) issue(blockIdx.x + i * gridDim.x, i);
  for (int i = 0; i < n; i++) {
    const int stage = i % S, t = blockIdx.x + i * gridDim.x;
    if (tid == 0 && i + S - 1 < n) {
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the stage of tile i-1 is free once its store has read smem
      issue(blockIdx.x + (i + S - 1) * gridDim.x, (i + S - 1) % S);
    }
    mbar_wait(&full[stage], (i / S) & 1);
    uint8_t* base = smem + stage * STAGE_BYTES;
    if (MODE == 1) {
      uint4* p = reinterpret_cast<uint4*>(base);
#pragma unroll
      for (int k = 0; k < STAGE_BYTES / 16 / 128; k++) { uint4 v = p[tid + k * 128]; v.x ^= 1; p[tid + k * 128] = v; }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < STAGE_BYTES / 16 / 128; k++) { uint4 v = p[((tid + 5) & 127) + k * 128]; v.y ^= 1; p[((tid + 5) & 127) + k * 128] = v; }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
      const int slot = t / per_slot, r = t % per_slot, ty = r / tiles_x, tx = r % tiles_x;
      tma_store3(&mp->dy, base, tx * 128 - 4, ty * 32 - 4, slot);
      tma_store3(&mp->dcb, base + 8192, tx * 64 - 4, ty * 16 - 2, slot);
      tma_store3(&mp->dcr, base + 8192 + 2048, tx * 64 - 4, ty * 16 - 2, slot);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  }
  if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                             CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static CUtensorMap make_map(EncodeFn enc, int16_t* base, int w, int h, int pitch, int bw, int bh) {
  CUtensorMap m;
  cuuint64_t dims[3] = {(cuuint64_t)w, (cuuint64_t)h, NS};
  cuuint64_t strides[2] = {(cuuint64_t)pitch * 2, (cuuint64_t)BUF * 2};
  cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 1};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  fprintf(stderr, "encode -> %d\n", (int)r);
  return m;
}

template <int S, int MODE>
void run(const Maps* m, int ctas_per_sm, double bytes, int16_t* src, int16_t* dst) {
  const int tiles_x = 31, tiles_y = 68, total = tiles_x * tiles_y * NS;
  const int smem = S * STAGE_BYTES + 8 * S;
  fprintf(stderr, "run S=%d mode=%d\n", S, MODE);
  cudaFuncSetAttribute(tma_copy<S, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  const int grid = 148 * ctas_per_sm;
  fprintf(stderr, "attr set, launching\n");
  for (int i = 0; i < 3; i++) tma_copy<S, MODE><<<grid, 128, smem>>>(m, tiles_x, tiles_y, total);
  fprintf(stderr, "warm launched: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  cudaEventRecord(a);
  const int N = 20;
  for (int i = 0; i < N; i++) tma_copy<S, MODE><<<grid, 128, smem>>>(m, tiles_x, tiles_y, total);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  // verify
  cudaError_t e = cudaGetLastError();
  printf("S=%d mode=%d ctas/SM=%d smem/SM=%dKB  %8.3f ms  %8.1f GB/s  %s\n", S, MODE, ctas_per_sm, ctas_per_sm * smem / 1024, ms / N, bytes / (ms / N * 1e-3) / 1e9, cudaGetErrorString(e));
}

int main() {
  int16_t *src, *dst;
  cudaMalloc(&src, NS * BUF * 2); cudaMalloc(&dst, NS * BUF * 2);
  int16_t* h = new int16_t[BUF];
  for (size_t i = 0; i < BUF; i++) h[i] = (int16_t)(i * 2654435761u >> 20);
  for (int s = 0; s < NS; s++) cudaMemcpy(src + s * BUF, h, BUF * 2, cudaMemcpyHostToDevice);
  cudaMemset(dst, 0, NS * BUF * 2);
  EncodeFn enc = nullptr;
  cudaDriverEntryPointQueryResult qres;
  fprintf(stderr, "alloc done\n");
  cudaError_t ge = cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", (void**)&enc, 12000, cudaEnableDefault, &qres);
  fprintf(stderr, "entry point: %s q=%d fn=%p\n", cudaGetErrorString(ge), (int)qres, (void*)enc);
  if (!enc) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
  Maps m;
  m.sy = make_map(enc, src, W, H, PY, 128, 32);
  m.scb = make_map(enc, src + PLANE_Y, W / 2, H / 2, PC, 64, 16);
  m.scr = make_map(enc, src + PLANE_Y + PLANE_C, W / 2, H / 2, PC, 64, 16);
  m.dy = make_map(enc, dst, W, H, PY, 128, 32);
  m.dcb = make_map(enc, dst + PLANE_Y, W / 2, H / 2, PC, 64, 16);
  m.dcr = make_map(enc, dst + PLANE_Y + PLANE_C, W / 2, H / 2, PC, 64, 16);
  const double bytes = 2.0 * NS * BUF * 2;
  Maps* dm; cudaMalloc(&dm, sizeof(Maps)); cudaMemcpy(dm, &m, sizeof(Maps), cudaMemcpyHostToDevice);
  fprintf(stderr, "maps done\n");
  run<2, 0>(dm, 8, bytes, src, dst);
  // correctness of the pure copy
  {
    int16_t* o = new int16_t[BUF];
    cudaMemcpy(o, dst + 3 * BUF, BUF * 2, cudaMemcpyDeviceToHost);
    size_t bad = 0;
    for (size_t i = 0; i < BUF; i++) bad += o[i] != h[i];
    printf("pure-TMA copy mismatches in slot 3: %zu\n", bad);
  }
  run<3, 0>(dm, 4, bytes, src, dst);
  run<4, 0>(dm, 4, bytes, src, dst);
  run<4, 0>(dm, 2, bytes, src, dst);
  run<8, 0>(dm, 2, bytes, src, dst);
  run<8, 0>(dm, 1, bytes, src, dst);
  run<16, 0>(dm, 1, bytes, src, dst);
  run<2, 1>(dm, 8, bytes, src, dst);
  run<3, 1>(dm, 4, bytes, src, dst);
  run<3, 1>(dm, 5, bytes, src, dst);
  run<4, 1>(dm, 4, bytes, src, dst);
  run<6, 1>(dm, 3, bytes, src, dst);
  run<8, 1>(dm, 2, bytes, src, dst);
  run<4, 1>(dm, 2, bytes, src, dst);
  return 0;
}
