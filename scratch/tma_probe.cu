#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <vector>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int VARIANT>
__global__ void probe(const __grid_constant__ CUtensorMap map, double* out, int x, int y, int z, int W, int R) {
  extern __shared__ __align__(128) unsigned char raw[];
  double* tile = (double*)raw;
  __shared__ __align__(8) uint64_t bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1) : "memory");
    if (VARIANT == 0) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    else asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(W * R * 8) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_u32(tile)), "l"(&map), "r"(smem_u32(&bar)), "r"(x), "r"(y), "r"(z) : "memory");
  }
  uint32_t done;
  do {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
  } while (!done);
  for (int t = threadIdx.x; t < W * R; t += blockDim.x) out[t] = tile[t];
}
typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int run(int variant, CUtensorMapDataType dt, int W, int R, int esz) {
  const int sx = 70, sy = 50, sz = 20;
  std::vector<double> h((size_t)sx * sy * sz);
  for (size_t t = 0; t < h.size(); ++t) h[t] = (double)t;
  double *d, *o; cudaMalloc(&d, h.size() * 8); cudaMalloc(&o, 8 * 4096);
  cudaMemcpy(d, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
  void* p; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  Enc enc = (Enc)p;
  CUtensorMap map;
  cuuint64_t dims[3] = {(cuuint64_t)(sx * 8 / esz), sy, sz};
  cuuint64_t strides[2] = {(cuuint64_t)sx * 8, (cuuint64_t)sx * sy * 8};
  cuuint32_t box[3] = {(cuuint32_t)(W * 8 / esz), (cuuint32_t)R, 1}, es[3] = {1, 1, 1};
  CUresult r = enc(&map, dt, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("variant %d dt %d W %d R %d: encode %d\n", variant, (int)dt, W, R, (int)r);
  int x = 3 * 8 / esz, y = 2, z = 1;
  if (variant == 0) probe<0><<<1, 128, W * R * 8 + 128>>>(map, o, x, y, z, W, R);
  else probe<1><<<1, 128, W * R * 8 + 128>>>(map, o, x, y, z, W, R);
  cudaError_t e = cudaDeviceSynchronize();
  printf("  sync: %s\n", cudaGetErrorString(e));
  if (e != cudaSuccess) return 1;
  std::vector<double> ho(W * R);
  cudaMemcpy(ho.data(), o, W * R * 8, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int rr = 0; rr < R; ++rr) for (int c = 0; c < W; ++c) {
    double want = (double)((size_t)(z) * sx * sy + (size_t)(y + rr) * sx + (3 + c));
    if (ho[rr * W + c] != want) ++bad;
  }
  printf("  mismatches: %d (first %.0f)\n", bad, ho[0]);
  return 0;
}
int main(int argc, char** argv) {
  int which = argc > 1 ? atoi(argv[1]) : 0;
  if (which == 0) return run(0, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 34, 14, 8);
  if (which == 1) return run(1, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 34, 14, 8);
  if (which == 2) return run(0, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 32, 14, 8);
  if (which == 3) return run(0, CU_TENSOR_MAP_DATA_TYPE_UINT64, 34, 14, 8);
  if (which == 4) return run(0, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 34, 14, 4);
  if (which == 5) return run(0, CU_TENSOR_MAP_DATA_TYPE_UINT32, 34, 14, 4);
  return 0;
}
