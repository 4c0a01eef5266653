// minimal TMA load/store bisect: ./tma_test <rank 2|3> <x> <y> <boxw> <boxh> <mapspace 0=global 1=param>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s2u(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int RANK>
__device__ void body(const CUtensorMap* ms, const CUtensorMap* md, int x, int y, int bytes) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s2u(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s2u(&bar)), "r"(bytes) : "memory");
    if (RANK == 2)
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(s2u(smem)), "l"(ms), "r"(s2u(&bar)), "r"(x), "r"(y) : "memory");
    else
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(s2u(smem)), "l"(ms), "r"(s2u(&bar)), "r"(x), "r"(y), "r"(1) : "memory");
  }
  asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}" ::"r"(s2u(&bar)), "r"(0) : "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    if (RANK == 2) asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(md), "r"(s2u(smem)), "r"(x), "r"(y) : "memory");
    else asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(md), "r"(s2u(smem)), "r"(x), "r"(y), "r"(1) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}
template <int RANK> __global__ void k_global(const CUtensorMap* ms, const CUtensorMap* md, int x, int y, int bytes) { body<RANK>(ms, md, x, y, bytes); }
template <int RANK> __global__ void k_param(const __grid_constant__ CUtensorMap ms, const __grid_constant__ CUtensorMap md, int x, int y, int bytes) { body<RANK>(&ms, &md, x, y, bytes); }
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                             CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main(int argc, char** argv) {
  const int rank = atoi(argv[1]), x = atoi(argv[2]), y = atoi(argv[3]), bw = atoi(argv[4]), bh = atoi(argv[5]), space = atoi(argv[6]);
  const int W = 3840, H = 2160, NS = 3;
  int16_t *src, *dst;
  const size_t plane = (size_t)W * H;
  cudaMalloc(&src, NS * plane * 2); cudaMalloc(&dst, NS * plane * 2);
  cudaMemset(src, 7, NS * plane * 2); cudaMemset(dst, 0, NS * plane * 2);
  EncodeFn enc = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", (void**)&enc, 12000, cudaEnableDefault, &q);
  CUtensorMap ms, md;
  cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)NS};
  cuuint64_t strides[2] = {(cuuint64_t)W * 2, (cuuint64_t)plane * 2};
  cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 1}, es[3] = {1, 1, 1};
  int r1 = enc(&ms, CU_TENSOR_MAP_DATA_TYPE_UINT16, rank, src, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  int r2 = enc(&md, CU_TENSOR_MAP_DATA_TYPE_UINT16, rank, dst, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CUtensorMap* dm; cudaMalloc(&dm, 256); cudaMemcpy(dm, &ms, 128, cudaMemcpyHostToDevice); cudaMemcpy(dm + 1, &md, 128, cudaMemcpyHostToDevice);
  const int bytes = bw * bh * 2;
  if (space == 0) { if (rank == 2) k_global<2><<<1, 128, bytes>>>(dm, dm + 1, x, y, bytes); else k_global<3><<<1, 128, bytes>>>(dm, dm + 1, x, y, bytes); }
  else { if (rank == 2) k_param<2><<<1, 128, bytes>>>(ms, md, x, y, bytes); else k_param<3><<<1, 128, bytes>>>(ms, md, x, y, bytes); }
  cudaError_t e = cudaDeviceSynchronize();
  int16_t v[4] = {0, 0, 0, 0};
  if (e == cudaSuccess) cudaMemcpy(v, dst + (rank == 3 ? plane : 0) + (size_t)(y < 0 ? 0 : y) * W + (x < 0 ? 0 : x), 8, cudaMemcpyDeviceToHost);
  printf("rank=%d x=%d y=%d box=%dx%d space=%d enc=%d,%d -> %s  dst[..]=%d %d\n", rank, x, y, bw, bh, space, r1, r2, cudaGetErrorString(e), v[0], v[3]);
  return 0;
}
