// ilf_ring.cuh -- TMA tile ring used by the band-walking kernels (sm_100a).
//
// A CTA owns a horizontal band of a plane and walks it left to right in tiles of TILE_W samples.  Tiles (plus halo
// rows) are fetched by the TMA unit (cp.async.bulk.tensor) into a ring of S shared-memory stages, each guarded by an
// mbarrier; thread 0 issues the loads up to S tiles ahead, so the bytes in flight per SM are set by the ring depth and
// not by the number of resident threads or their registers.  All boxes start at an x that is a multiple of TILE_W
// (16-byte aligned, no horizontal over-fetch): the horizontal halo of a tile is simply the neighbouring tile, which is
// in the ring anyway.  Out-of-picture parts of a box are zero-filled by the TMA unit.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace ilf {
namespace ring {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}
// box of a 3-D tensor (x, y, z) -> shared memory; completion is signalled on `bar` with the box's byte count
// ILF_TMA_EVICT_FIRST=1: picture data is streamed once per stage, so the loads carry an evict-first L2 policy.
#ifndef ILF_TMA_EVICT_FIRST
#define ILF_TMA_EVICT_FIRST 0
#endif
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int x, int y, int z) {
#if ILF_TMA_EVICT_FIRST
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5}], [%2], %6;" ::"r"(smem_u32(dst)),
               "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z), "l"(pol)
               : "memory");
#else
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(dst)), "l"(map),
               "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z)
               : "memory");
#endif
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int x, int y) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)), "l"(map),
               "r"(smem_u32(bar)), "r"(x), "r"(y)
               : "memory");
}

// Ring bookkeeping for a walk over tiles [first, last] (inclusive): tile t lives in stage (t - first) % S and completes
// phase ((t - first) / S) & 1 of that stage's barrier.
template <int S>
struct Walk {
  int first, last;
  __device__ __forceinline__ int stage(int t) const { return (t - first) % S; }
  __device__ __forceinline__ uint32_t parity(int t) const { return (uint32_t)(((t - first) / S) & 1); }
};

}  // namespace ring
}  // namespace ilf
