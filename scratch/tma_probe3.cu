#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda/barrier>
#include <stdio.h>
#include <stdint.h>
#include <vector>
using barrier = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int RANK>
__global__ void k_api(const __grid_constant__ CUtensorMap map, double* out, int x, int y, int z, int n) {
  extern __shared__ __align__(128) unsigned char raw[];
  double* tile = (double*)raw;
  __shared__ barrier bar;
  if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
  __syncthreads();
  barrier::arrival_token token;
  if (threadIdx.x == 0) {
    if (RANK == 2) cde::cp_async_bulk_tensor_2d_global_to_shared(tile, &map, x, y, bar);
    else cde::cp_async_bulk_tensor_3d_global_to_shared(tile, &map, x, y, z, bar);
    token = cuda::device::barrier_arrive_tx(bar, 1, n * 8);
  } else token = bar.arrive();
  bar.wait(std::move(token));
  for (int t = threadIdx.x; t < n; t += blockDim.x) out[t] = tile[t];
}
template <int RANK>
__global__ void k_asm(const __grid_constant__ CUtensorMap map, double* out, int x, int y, int z, int n) {
  extern __shared__ __align__(128) unsigned char raw[];
  double* tile = (double*)raw;
  __shared__ __align__(8) uint64_t bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1) : "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(n * 8) : "memory");
    if (RANK == 2)
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                   ::"r"(smem_u32(tile)), "l"(&map), "r"(smem_u32(&bar)), "r"(x), "r"(y) : "memory");
    else
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                   ::"r"(smem_u32(tile)), "l"(&map), "r"(smem_u32(&bar)), "r"(x), "r"(y), "r"(z) : "memory");
  }
  uint32_t done;
  do {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
  } while (!done);
  for (int t = threadIdx.x; t < n; t += blockDim.x) out[t] = tile[t];
}
typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main(int argc, char** argv) {
  int which = argc > 1 ? atoi(argv[1]) : 0;
  const int sx = 64, sy = 50, sz = 20;
  std::vector<double> h((size_t)sx * sy * sz);
  for (size_t t = 0; t < h.size(); ++t) h[t] = (double)t;
  double *d, *o; cudaMalloc(&d, h.size() * 8); cudaMalloc(&o, 8 * 4096);
  cudaMemcpy(d, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
  void* p; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  Enc enc = (Enc)p;
  struct Cfg { int api, rank, W, R; } cfgs[] = {{1, 3, 32, 16}, {0, 2, 32, 16}, {0, 3, 32, 16}, {1, 3, 34, 14}, {1, 2, 34, 14}, {0, 2, 34, 14}};
  Cfg c = cfgs[which];
  CUtensorMap map;
  CUresult r;
  if (c.rank == 2) {
    cuuint64_t dims[2] = {sx, (cuuint64_t)sy * sz}; cuuint64_t strides[1] = {(cuuint64_t)sx * 8};
    cuuint32_t box[2] = {(cuuint32_t)c.W, (cuuint32_t)c.R}, es[2] = {1, 1};
    r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else {
    cuuint64_t dims[3] = {sx, sy, sz}; cuuint64_t strides[2] = {(cuuint64_t)sx * 8, (cuuint64_t)sx * sy * 8};
    cuuint32_t box[3] = {(cuuint32_t)c.W, (cuuint32_t)c.R, 1}, es[3] = {1, 1, 1};
    r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  printf("cfg %d api=%d rank=%d W=%d R=%d encode %d\n", which, c.api, c.rank, c.W, c.R, (int)r);
  int n = c.W * c.R;
  size_t sm = n * 8 + 128;
  if (c.api) { if (c.rank == 2) k_api<2><<<1, 128, sm>>>(map, o, 4, 2, 1, n); else k_api<3><<<1, 128, sm>>>(map, o, 4, 2, 1, n); }
  else { if (c.rank == 2) k_asm<2><<<1, 128, sm>>>(map, o, 4, 2, 1, n); else k_asm<3><<<1, 128, sm>>>(map, o, 4, 2, 1, n); }
  printf("  sync: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  double v[2]; cudaMemcpy(v, o, 16, cudaMemcpyDeviceToHost);
  printf("  first %.0f (want %d)\n", v[0], (c.rank == 3 ? sx * sy : 0) + 2 * sx + 4);
  return 0;
}
