// mbarrier / TMA helpers and the tensor-map encoder shared by the pipelined kernels (ocb_sweep_fused.cu, ocb_sweep_strain.cu)
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include "ocb_common.cuh"

namespace {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
  uint32_t done;
  const uint32_t addr = smem_u32(bar);
  do {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(done) : "r"(addr), "r"(phase) : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int x, int y, int z) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z) : "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = (EncodeTiledFn)p;
  return fn;
}


// Float64 fields: dense parents whose row pitch is a 16-byte multiple
static bool tma_compatible_f64(const ocb_field& f, const ocb_grid* g) {
  if (f.kind != OCB_FIELD_ARRAY || !f.data) return false;
  for (int d = 0; d < 3; ++d) if (f.loc[d] == OCB_NOTHING || f.halo[d] != g->H[d]) return false;
  if (f.stride[0] != 1) return false;
  if (((uintptr_t)f.data & 15) || ((f.stride[1] * 8) & 15) || ((f.stride[2] * 8) & 15)) return false;
  if (f.stride[1] != f.size[0] || f.stride[2] != (long long)f.size[0] * f.size[1]) return false;   // dense parent
  return true;
}
static int make_map_f64(EncodeTiledFn enc, const ocb_field& f, int rows, CUtensorMap* map, int tw) {
  cuuint64_t dims[3] = {(cuuint64_t)f.size[0], (cuuint64_t)f.size[1], (cuuint64_t)f.size[2]};
  cuuint64_t strides[2] = {(cuuint64_t)f.stride[1] * 8, (cuuint64_t)f.stride[2] * 8};
  cuuint32_t box[3] = {(cuuint32_t)tw, (cuuint32_t)rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, f.data, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) OCB_FAIL(OCB_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return OCB_OK;
}

// Float32 fields: the merged-row view (two parent rows per tensor row) must stride by 16-byte multiples
static bool tma_compatible_f32(const ocb_field& f, const ocb_grid* g) {
  if (f.kind != OCB_FIELD_ARRAY || !f.data) return false;
  for (int d = 0; d < 3; ++d) if (f.loc[d] == OCB_NOTHING || f.halo[d] != g->H[d]) return false;
  if (f.stride[0] != 1) return false;
  if (f.stride[1] != f.size[0] || f.stride[2] != (long long)f.size[0] * f.size[1]) return false;   // dense parent
  if (((uintptr_t)f.data & 15) || (f.size[0] & 1) || (f.size[1] & 1) || ((f.stride[2] * 4) & 15)) return false;
  return true;
}
static int make_map_f32(EncodeTiledFn enc, const ocb_field& f, int rows, CUtensorMap* map, int tw) {
  cuuint64_t dims[3] = {(cuuint64_t)f.size[0] * 2, (cuuint64_t)f.size[1] / 2, (cuuint64_t)f.size[2]};
  cuuint64_t strides[2] = {(cuuint64_t)f.stride[1] * 2 * 4, (cuuint64_t)f.stride[2] * 4};
  cuuint32_t box[3] = {(cuuint32_t)tw, (cuuint32_t)(rows / 2), 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, f.data, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) OCB_FAIL(OCB_ERR_CUDA, "cuTensorMapEncodeTiled (Float32, merged rows) failed with CUresult %d", (int)r);
  return OCB_OK;
}

}  // namespace
