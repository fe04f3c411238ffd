// Library-level entry points: ABI version, last-error string, device probing, and the shared
// deterministic final reduction used by the fused volume integrals.
#include "ocb_common.cuh"

static thread_local char g_err[1024] = "";

void ocb_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" int ocb_abi_version(void) { return OCB_ABI_VERSION; }
extern "C" const char* ocb_last_error(void) { return g_err; }

extern "C" int ocb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int ocb_require_device() {
  if (ocb_device_count() <= 0)
    OCB_FAIL(OCB_ERR_CUDA, "no usable CUDA device: liboceanostics_b200 has no CPU fallback");
  return OCB_OK;
}

int ocb_sm_count() {
  static thread_local int cached_dev = -1, cached = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    cudaDeviceGetAttribute(&cached, cudaDevAttrMultiProcessorCount, dev);
    cached_dev = dev;
  }
  return cached > 0 ? cached : 148;
}

// Stream-ordered scratch (cudaMallocAsync) comes from the device's default pool; by default the pool hands memory
// back to the OS at every synchronisation, which would make each filter / sort call re-allocate gigabytes.  Keep it.
void ocb_keep_scratch_pool() {
  static thread_local int done_dev = -1;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev == done_dev) return;
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
    unsigned long long keep = ~0ull;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
  }
  done_dev = dev;
}

// One block per integral slot sums the per-block partials in a fixed order (thread t takes blocks
// t, t+256, ...; then a fixed-shape tree), so a plan gives the same bits run to run.
__global__ void __launch_bounds__(256) ocb_finalize_partials_kernel(const double* __restrict__ partial, int nblocks,
                                                                     int nslots, double* __restrict__ out) {
  __shared__ double sh[256];
  const int s = blockIdx.x;
  double acc = 0.0;
  for (int b = threadIdx.x; b < nblocks; b += 256) acc += partial[(size_t)b * nslots + s];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if (threadIdx.x < w) sh[threadIdx.x] += sh[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[s] = sh[0];
}

void ocb_launch_finalize_partials(const double* partial, int nblocks, int nslots, double* out, cudaStream_t st) {
  ocb_finalize_partials_kernel<<<nslots, 256, 0, st>>>(partial, nblocks, nslots, out);
}
