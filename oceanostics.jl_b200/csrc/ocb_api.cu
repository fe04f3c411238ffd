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

long long ocb_cycles(double seconds) {
  // cudaDevAttrClockRate is one of the attributes the runtime fetches from the device on every query (slow, and serialised
  // between the processes of a box): asked once per device, never on a per-step path
  static int khz_of[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
  if (khz_of[dev] <= 0) {
    int khz = 0;
    if (cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev) != cudaSuccess || khz <= 0) khz = 1965000;
    khz_of[dev] = khz;
  }
  return (long long)(seconds * 1e3 * (double)khz_of[dev]);
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

// ---- the same, followed by the sum over the ranks of a peer-memory ring (ocb_plan_execute_ring) -----------------------
// One block.  Slot sums exactly as above (same order, same bits); thread r then stores them into rank r's sync block and
// publishes the step number behind a system-scope fence; threads r < nranks wait for rank r's contribution to this
// rank's block and thread 0 adds them in rank order.  Two parities: a rank can be one step ahead of a slow reader, never two
// (it needs everyone's contribution of step k+1, sent after their step-k reads, before it can send step k+2).
struct RingPeers { long long* sync[OCB_MAX_RANKS]; };
__global__ void __launch_bounds__(256) ocb_finalize_ring_kernel(const double* __restrict__ partial, int nblocks, int nslots,
                                                                 double* __restrict__ out, const __grid_constant__ RingPeers peers,
                                                                 int rank, int nranks, long long step, long long timeout_cycles) {
  __shared__ double sh[256];
  __shared__ double mine[16];   // OCB_MAX_REQ slots at most
  __shared__ int timed_out;
  const int t = threadIdx.x;
  if (t == 0) timed_out = 0;
  for (int s = 0; s < nslots; ++s) {
    double acc = 0.0;
    for (int b = t; b < nblocks; b += 256) acc += partial[(size_t)b * nslots + s];
    sh[t] = acc;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
      if (t < w) sh[t] += sh[t + w];
      __syncthreads();
    }
    if (t == 0) mine[s] = sh[0];
    __syncthreads();
  }
  const int par = (int)(step & 1);
  if (t < nranks) {
    long long* dst = peers.sync[t];
    double* slots = reinterpret_cast<double*>(dst + 64 + 16 * (16 * par + rank));
    for (int s = 0; s < nslots; ++s) slots[s] = mine[s];
    __threadfence_system();
    asm volatile("st.release.sys.global.s64 [%0], %1;" ::"l"(dst + 8 + 16 * par + rank), "l"(step) : "memory");
  }
  long long* own = peers.sync[rank];
  if (t < nranks) {
    const long long t0 = clock64();
    for (;;) {
      long long seen;
      asm volatile("ld.acquire.sys.global.s64 %0, [%1];" : "=l"(seen) : "l"(own + 8 + 16 * par + t) : "memory");
      if (seen >= step) break;
      if (clock64() - t0 > timeout_cycles) { timed_out = 1; own[2] = 1; break; }
      __nanosleep(50);
    }
  }
  __syncthreads();
  if (t < nslots) {
    double total = 0.0;
    for (int r = 0; r < nranks; ++r)
      total += *reinterpret_cast<volatile double*>(own + 64 + 16 * (16 * par + r) + t);
    out[t] = timed_out ? __longlong_as_double(0x7ff8000000000000LL) : total;
  }
}

int ocb_launch_finalize_ring(const double* partial, int nblocks, int nslots, double* out, const ocb_ring* ring, long long step,
                             long long timeout_cycles, cudaStream_t st) {
  RingPeers P;
  for (int r = 0; r < OCB_MAX_RANKS; ++r) P.sync[r] = r < ring->nranks ? (long long*)ring->sync[r] : nullptr;
  ocb_finalize_ring_kernel<<<1, 256, 0, st>>>(partial, nblocks, nslots, out, P, ring->rank, ring->nranks, step, timeout_cycles);
  OCB_CUDA(cudaGetLastError());
  return OCB_OK;
}
