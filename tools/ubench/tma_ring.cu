// tma_ring.cu -- skeleton of the planned deblocking kernel: persistent CTAs, ring of S stages filled by TMA loads
// (16-byte aligned superset boxes: luma 136x32 at (x0-8, y0-4), chroma 72x16 at (cx0-8, cy0-2), metadata 36x8 units),
// tile (x0-4.., y0-4..) written back by the threads with 8-byte stores.  MODE 1 adds a read-modify-write pass + 2 syncs.
// MODE 2: the same ring filled by the threads with 16-byte cp.async instead of TMA boxes (commit / wait groups instead of mbarriers):
// separates "TMA" from "persistent CTAs walking a ring" as the reason why the skeleton stays below a flat copy.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

constexpr int W = 3840, H = 2160, NS = 17;
constexpr int PY = 3840, PC = 1920, UW = 960, UH = 540;
constexpr size_t PLANE_Y = (size_t)PY * H, PLANE_C = (size_t)PC * (H / 2), BUF = PLANE_Y + 2 * PLANE_C;
struct Cfg { int LP, CP, XO, CXO, META, S, OFF_C, OFF_INFO, OFF_MV, STAGE_BYTES, TX_BYTES, ORDER, STA; };   // STA: the box carries halo columns, the ALIGNED tile is stored
static Cfg make_cfg(int lp, int cp, int xo, int cxo, int meta, int S) {
  Cfg c; c.ORDER = 0; c.STA = 0; c.LP = lp; c.CP = cp; c.XO = xo; c.CXO = cxo; c.META = meta; c.S = S;
  c.OFF_C = lp * 32 * 2; c.OFF_INFO = c.OFF_C + 2 * cp * 16 * 2; c.OFF_MV = c.OFF_INFO + 36 * 8 * 4;
  c.TX_BYTES = meta ? c.OFF_MV + 36 * 8 * 8 : c.OFF_INFO;
  c.STAGE_BYTES = (c.TX_BYTES + 127) & ~127;
  return c;
}

__device__ __forceinline__ uint32_t s2u(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s2u(b)), "r"(n)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s2u(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}" ::"r"(s2u(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load3(void* dst, const CUtensorMap* m, uint64_t* bar, int x, int y, int z) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(s2u(dst)), "l"(m), "r"(s2u(bar)),
               "r"(x), "r"(y), "r"(z) : "memory");
}
struct Maps { CUtensorMap y, cb, cr, info, mv; };

__device__ __forceinline__ void cp16(void* dst, const void* src, bool valid) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s2u(dst)), "l"(src), "r"(valid ? 16 : 0) : "memory");
}
template <int MODE>
__global__ void __launch_bounds__(128) ring(const Maps* __restrict__ mp, int16_t* __restrict__ dst, int tiles_x, int tiles_y, int total, const Cfg c, const int16_t* __restrict__ srcp) {
  const int S = c.S, STAGE_BYTES = c.STAGE_BYTES, TX_BYTES = c.TX_BYTES, OFF_C = c.OFF_C, OFF_INFO = c.OFF_INFO, OFF_MV = c.OFF_MV, LP = c.LP, CP = c.CP;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + S * STAGE_BYTES);
  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int i = 0; i < S; i++) mbar_init(&full[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int per_slot = tiles_x * tiles_y;
  // ORDER 0: tiles in raster order, strided over the CTAs; 1: a CTA walks a band (row of tiles) left to right, bands dealt
  // round-robin; 2: a CTA walks a column of tiles top to bottom
  auto decode = [&](int i, int& slot, int& ty, int& tx) {
    if (c.ORDER == 0) { const int t = blockIdx.x + i * gridDim.x; slot = t / per_slot; const int r = t % per_slot; ty = r / tiles_x; tx = r % tiles_x; }
    else if (c.ORDER == 1) { const int b = blockIdx.x + (i / tiles_x) * gridDim.x; slot = b / tiles_y; ty = b % tiles_y; tx = i % tiles_x; }
    else { const int b = blockIdx.x + (i / tiles_y) * gridDim.x; slot = b / tiles_x; tx = b % tiles_x; ty = i % tiles_y; }
  };
  auto issue = [&](int i, int stage) {
    int slot, ty, tx; decode(i, slot, ty, tx);
    uint8_t* base = smem + stage * STAGE_BYTES;
    mbar_expect(&full[stage], TX_BYTES);
    tma_load3(base, &mp->y, &full[stage], tx * 128 + c.XO, ty * 32 - 4, slot);
    tma_load3(base + OFF_C, &mp->cb, &full[stage], tx * 64 + c.CXO, ty * 16 - 2, slot);
    tma_load3(base + OFF_C + CP * 16 * 2, &mp->cr, &full[stage], tx * 64 + c.CXO, ty * 16 - 2, slot);
    if (c.META) {
    tma_load3(base + OFF_INFO, &mp->info, &full[stage], tx * 32 - 4, ty * 8 - 1, slot);
    tma_load3(base + OFF_MV, &mp->mv, &full[stage], (tx * 32 - 4) * 2, ty * 8 - 1, slot);
    }
  };
  // MODE 2: every thread copies 4 luma + 2 chroma 16-byte chunks of the tile (box 128 x 32 at (x, y - 4), 2 x 64 x 16 at (x / 2, y / 2 - 2))
  auto issue_cp = [&](int i, int stage) {
    int slot, ty, tx; decode(i, slot, ty, tx);
    uint8_t* base = smem + stage * STAGE_BYTES;
    const int16_t* sp = srcp + (size_t)slot * BUF;
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const int ch = tid + q * 128, row = ch >> 4, cx = ch & 15, y = ty * 32 - 4 + row;
      const bool ok = y >= 0 && y < H;
      cp16(base + (row * LP + 8 * cx) * 2, sp + (size_t)(ok ? y : 0) * PY + tx * 128 + 8 * cx, ok);
    }
#pragma unroll
    for (int q = 0; q < 2; q++) {
      const int ch = tid + q * 128, pl = ch >> 7, row = (ch >> 3) & 15, cx = ch & 7, y = ty * 16 - 2 + row;
      const bool ok = y >= 0 && y < H / 2;
      cp16(base + OFF_C + (pl * CP * 16 + row * CP + 8 * cx) * 2, sp + PLANE_Y + pl * PLANE_C + (size_t)(ok ? y : 0) * PC + tx * 64 + 8 * cx, ok);
    }
  };
  int n = 0;
  if (c.ORDER == 0) { for (int t = blockIdx.x; t < total; t += gridDim.x) n++; }
  else if (c.ORDER == 1) { for (int b = blockIdx.x; b < tiles_y * NS; b += gridDim.x) n += tiles_x; }
  else { for (int b = blockIdx.x; b < tiles_x * NS; b += gridDim.x) n += tiles_y; }
  if (MODE == 2) {
    for (int i = 0; i < S - 1; i++) { if (i < n) issue_cp(i, i); asm volatile("cp.async.commit_group;" ::: "memory"); }
  } else if (tid == 0)
    for (int i = 0; i < S && i < n; i++) issue(i, i);
  const int k = tid & 31, r0 = tid >> 5, kc = tid & 15, rc0 = tid >> 4;
  for (int i = 0; i < n; i++) {
    const int stage = i % S;
    int slot, ty, tx; decode(i, slot, ty, tx);
    if (MODE == 2) {
      if (S == 2) asm volatile("cp.async.wait_group 0;" ::: "memory");
      else if (S == 3) asm volatile("cp.async.wait_group 1;" ::: "memory");
      else asm volatile("cp.async.wait_group 2;" ::: "memory");
      __syncthreads();   // the tile has landed for every thread; every thread is done with the stage refilled below
      if (i + S - 1 < n) issue_cp(i + S - 1, (i + S - 1) % S);
      asm volatile("cp.async.commit_group;" ::: "memory");
    } else mbar_wait(&full[stage], (i / S) & 1);
    uint8_t* base = smem + stage * STAGE_BYTES;
    int16_t* sy = reinterpret_cast<int16_t*>(base);
    int16_t* sc = reinterpret_cast<int16_t*>(base + OFF_C);
    if (MODE == 1) {
      uint4* p = reinterpret_cast<uint4*>(base);
#pragma unroll
      for (int q = 0; q < 8; q++) { uint4 v = p[tid + q * 128]; v.x ^= 1; p[tid + q * 128] = v; }
      __syncthreads();
#pragma unroll
      for (int q = 0; q < 8; q++) { uint4 v = p[((tid + 5) & 127) + q * 128]; v.y ^= 1; p[((tid + 5) & 127) + q * 128] = v; }
      __syncthreads();
    }
    int16_t* d = dst + (size_t)slot * BUF;
    const int sh = (c.XO && !c.STA) ? 4 : 0; const int x0 = tx * 128 - sh, y0 = ty * 32 - 4, cx0 = tx * 64 - sh, cy0 = ty * 16 - 2; const int so = c.XO ? -c.XO - sh : 0, sco = c.CXO ? -c.CXO - sh : 0;
#pragma unroll
    for (int q = 0; q < 8; q++) {
      const int x = x0 + 4 * k, y = y0 + r0 + 4 * q;
      if (x >= 0 && x < W && y >= 0 && y < H) *(uint2*)(d + (size_t)y * PY + x) = *(uint2*)&sy[(r0 + 4 * q) * LP + so + 4 * k];
    }
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const int pl = q >> 1, x = cx0 + 4 * kc, y = cy0 + rc0 + 8 * (q & 1);
      if (x >= 0 && x < W / 2 && y >= 0 && y < H / 2) *(uint2*)(d + PLANE_Y + pl * PLANE_C + (size_t)y * PC + x) = *(uint2*)&sc[pl * CP * 16 + (rc0 + 8 * (q & 1)) * CP + sco + 4 * kc];
    }
    if (MODE != 2) {
      __syncthreads();  // every thread is done reading the stage
      if (tid == 0 && i + S < n) issue(i + S, stage);
    }
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                             CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static CUtensorMapL2promotion g_l2x = CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
static CUtensorMap make_map(EncodeFn enc, CUtensorMapDataType dt, int es_bytes, void* base, int w, int h, size_t pitch_bytes, size_t slot_bytes, int bw, int bh) {
  CUtensorMap m;
  cuuint64_t dims[3] = {(cuuint64_t)w, (cuuint64_t)h, NS};
  cuuint64_t strides[2] = {(cuuint64_t)pitch_bytes, (cuuint64_t)slot_bytes};
  cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 1}, es[3] = {1, 1, 1};
  CUresult r = enc(&m, dt, 3, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, g_l2x, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) fprintf(stderr, "encode failed %d\n", (int)r);
  (void)es_bytes;
  return m;
}

int16_t *src, *dst; int16_t* h;
static EncodeFn g_enc; static uint32_t* g_info; static uint2* g_mv; static CUtensorMapL2promotion g_l2 = CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
template <int MODE>
void run(int lp, int cp, int xo, int cxo, int meta, int S, int ctas_per_sm, double bytes, int order = 0, int sta = 0) {
  Cfg c = make_cfg(lp, cp, xo, cxo, meta, S); c.ORDER = order; c.STA = sta;
  const int STAGE_BYTES = c.STAGE_BYTES;
  Maps hm;
  hm.y = make_map(g_enc, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, src, W, H, PY * 2, BUF * 2, lp, 32);
  hm.cb = make_map(g_enc, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, src + PLANE_Y, W / 2, H / 2, PC * 2, BUF * 2, cp, 16);
  hm.cr = make_map(g_enc, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, src + PLANE_Y + PLANE_C, W / 2, H / 2, PC * 2, BUF * 2, cp, 16);
  hm.info = make_map(g_enc, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, g_info, UW, UH, UW * 4, (size_t)UW * UH * 4, 36, 8);
  hm.mv = make_map(g_enc, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, g_mv, UW * 2, UH, UW * 8, (size_t)UW * UH * 8, 72, 8);
  Maps* m; cudaMalloc(&m, sizeof(Maps)); cudaMemcpy(m, &hm, sizeof(Maps), cudaMemcpyHostToDevice);
  const int tiles_x = (xo && !sta) ? 31 : 30, tiles_y = 68, total = tiles_x * tiles_y * NS;
  const int smem = S * STAGE_BYTES + 8 * S;
  cudaFuncSetAttribute(ring<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  const int grid = ctas_per_sm > 0 ? 148 * ctas_per_sm : total / -ctas_per_sm;   // negative: -k tiles per CTA (non-persistent CTAs)
  cudaMemset(dst, 0, NS * BUF * 2);
  for (int i = 0; i < 3; i++) ring<MODE><<<grid, 128, smem>>>(m, dst, tiles_x, tiles_y, total, c, src);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("S=%d mode=%d ctas/SM=%d: %s\n", S, MODE, ctas_per_sm, cudaGetErrorString(e)); exit(1); }
  cudaEventRecord(a);
  const int N = 20;
  for (int i = 0; i < N; i++) ring<MODE><<<grid, 128, smem>>>(m, dst, tiles_x, tiles_y, total, c, src);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  size_t bad = 0;
  if (MODE == 0 || MODE == 2) {
    int16_t* o = new int16_t[BUF];
    cudaMemcpy(o, dst + 5 * BUF, BUF * 2, cudaMemcpyDeviceToHost);
    for (size_t i = 0; i < BUF; i++) bad += o[i] != h[i];
    delete[] o;
  }
  int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ring<MODE>, 128, smem); cudaFree(m);
  printf("order=%d lp=%d cp=%d xo=%d sta=%d meta=%d S=%d mode=%d ctas/SM=%d (occ %d) smem/SM=%dKB  %8.3f ms  %8.1f GB/s (algorithmic, samples only)  mismatches=%zu\n", order, lp, cp, xo, sta, meta, S, MODE, ctas_per_sm, occ, ctas_per_sm * smem / 1024, ms / N,
         bytes / (ms / N * 1e-3) / 1e9, bad);
}

int main() {
  cudaMalloc(&src, NS * BUF * 2); cudaMalloc(&dst, NS * BUF * 2);
  h = new int16_t[BUF];
  for (size_t i = 0; i < BUF; i++) h[i] = (int16_t)(i * 2654435761u >> 20);
  for (int s = 0; s < NS; s++) cudaMemcpy(src + s * BUF, h, BUF * 2, cudaMemcpyHostToDevice);
  uint32_t* info; uint2* mv;
  cudaMalloc(&info, (size_t)NS * UW * UH * 4); cudaMalloc(&mv, (size_t)NS * UW * UH * 8);
  cudaMemset(info, 1, (size_t)NS * UW * UH * 4); cudaMemset(mv, 2, (size_t)NS * UW * UH * 8);
  EncodeFn enc = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", (void**)&enc, 12000, cudaEnableDefault, &q);
  g_enc = enc; g_info = info; g_mv = mv;
  const double bytes = 2.0 * NS * BUF * 2;
  if (getenv("RING_HALO_TILE")) {  // independent tiles with halo columns in the box (144 / 80 wide from x - 8), aligned stores; k tiles per CTA
    for (int meta = 0; meta < 2; meta++)
      for (int k = 1; k <= 6; k++) {
        if (k == 5) continue;
        for (int S = (k == 1 ? 1 : 2); S <= (k == 1 ? 1 : 3); S++) run<0>(144, 80, -8, -8, meta, S, -k, bytes, 0, 1);
      }
    for (int meta = 0; meta < 2; meta++) { run<0>(144, 80, -8, -8, meta, 3, 3, bytes, 0, 1); run<0>(144, 80, -8, -8, meta, 3, 3, bytes, 1, 1); }   // persistent for comparison
    return 0;
  }
  if (getenv("RING_ONE_TILE")) {  // non-persistent: one / two / four tiles per CTA, the hardware scheduler deals the CTAs
    for (int order = 0; order < 2; order++)
      for (int k = 1; k <= 4; k *= 2) {
        run<0>(128, 64, 0, 0, 0, k == 1 ? 1 : 2, -k, bytes, order);
        run<2>(128, 64, 0, 0, 0, 2, -k, bytes, order);
      }
    return 0;
  }
  if (getenv("RING_CP_ONLY")) {   // TMA against cp.async at the same ring depth / residency
    for (int order = 0; order < 2; order++)
      for (int ctas = 2; ctas <= 8; ctas *= 2)
        for (int S = 2; S <= 4; S++) {
          if (ctas * S * 12288 > 220 * 1024) continue;
          run<0>(128, 64, 0, 0, 0, S, ctas, bytes, order);
          run<2>(128, 64, 0, 0, 0, S, ctas, bytes, order);
        }
    return 0;
  }
  for (int order = 0; order < 3; order++) {
    run<0>(128, 64, 0, 0, 0, 2, 4, bytes, order);
    run<0>(128, 64, 0, 0, 0, 4, 4, bytes, order);
    run<0>(128, 64, 0, 0, 0, 3, 2, bytes, order);
    run<0>(128, 64, 0, 0, 1, 3, 4, bytes, order);
    run<0>(144, 80, -8, -8, 0, 3, 4, bytes, order);
  }
  return 0;
}
