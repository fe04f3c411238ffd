#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda/barrier>
#include <stdio.h>
#include <stdint.h>
#include <vector>
using barrier = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// (A) documented libcu++ path, 2D
__global__ void k_api2d(const __grid_constant__ CUtensorMap map, double* out, int x, int y) {
  __shared__ alignas(128) double tile[16][32];
  __shared__ barrier bar;
  if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
  __syncthreads();
  barrier::arrival_token token;
  if (threadIdx.x == 0) {
    cde::cp_async_bulk_tensor_2d_global_to_shared(&tile, &map, x, y, bar);
    token = cuda::device::barrier_arrive_tx(bar, 1, sizeof(tile));
  } else token = bar.arrive();
  bar.wait(std::move(token));
  for (int t = threadIdx.x; t < 16 * 32; t += blockDim.x) out[t] = (&tile[0][0])[t];
}
// (B) 1D bulk copy (no descriptor)
__global__ void k_bulk1d(const double* src, double* out, int n) {
  extern __shared__ __align__(128) unsigned char raw[];
  double* tile = (double*)raw;
  __shared__ __align__(8) uint64_t bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(n * 8) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(tile)), "l"(src), "r"(n * 8), "r"(smem_u32(&bar)) : "memory");
  }
  uint32_t done;
  do {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
  } while (!done);
  for (int t = threadIdx.x; t < n; t += blockDim.x) out[t] = tile[t];
}
// (C) descriptor copied to global memory instead of a kernel parameter
__global__ void k_gdesc(const CUtensorMap* map, double* out, int x, int y, int z, int W, int R) {
  extern __shared__ __align__(128) unsigned char raw[];
  double* tile = (double*)raw;
  __shared__ __align__(8) uint64_t bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1) : "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(W * R * 8) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_u32(tile)), "l"(map), "r"(smem_u32(&bar)), "r"(x), "r"(y), "r"(z) : "memory");
  }
  uint32_t done;
  do {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
  } while (!done);
  for (int t = threadIdx.x; t < W * R; t += blockDim.x) out[t] = tile[t];
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
  cudaError_t ge = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  printf("entry point: %s q=%d p=%p\n", cudaGetErrorString(ge), (int)q, p);
  Enc enc = (Enc)p;
  int drv = 0, rt = 0; cudaDriverGetVersion(&drv); cudaRuntimeGetVersion(&rt); printf("driver %d runtime %d\n", drv, rt);
  if (which == 0) {
    CUtensorMap map;
    cuuint64_t dims[2] = {sx, (cuuint64_t)sy * sz}; cuuint64_t strides[1] = {(cuuint64_t)sx * 8};
    cuuint32_t box[2] = {32, 16}, es[2] = {1, 1};
    CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("api2d encode %d\n", (int)r);
    k_api2d<<<1, 128>>>(map, o, 4, 2);
    printf("  api2d sync: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    double v[2]; cudaMemcpy(v, o, 16, cudaMemcpyDeviceToHost); printf("  first %.0f (want %d)\n", v[0], 2 * sx + 4);
  } else if (which == 1) {
    k_bulk1d<<<1, 128, 36 * 8 + 128>>>(d + 6, o, 36);
    printf("  bulk1d sync: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    double v[2]; cudaMemcpy(v, o, 16, cudaMemcpyDeviceToHost); printf("  first %.0f (want 6)\n", v[0]);
  } else {
    CUtensorMap map, *dmap;
    cuuint64_t dims[3] = {sx, sy, sz}; cuuint64_t strides[2] = {(cuuint64_t)sx * 8, (cuuint64_t)sx * sy * 8};
    cuuint32_t box[3] = {34, 14, 1}, es[3] = {1, 1, 1};
    CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("gdesc encode %d\n", (int)r);
    cudaMalloc(&dmap, sizeof(map)); cudaMemcpy(dmap, &map, sizeof(map), cudaMemcpyHostToDevice);
    k_gdesc<<<1, 128, 34 * 14 * 8 + 128>>>(dmap, o, 3, 2, 1, 34, 14);
    printf("  gdesc sync: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    double v[2]; cudaMemcpy(v, o, 16, cudaMemcpyDeviceToHost); printf("  first %.0f (want %d)\n", v[0], sx * sy + 2 * sx + 3);
  }
  return 0;
}
