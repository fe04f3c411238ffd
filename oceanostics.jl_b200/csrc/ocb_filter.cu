// K2: separable box / Gaussian spatial filters, staged as 1-D passes (reference: _compute_staged_filter!,
// src/SpatialFilters/SpatialFilters.jl:347-378; kernels box_filter.jl:19-47, gaussian_filter.jl:44-81 and :193-262;
// boundary policies SpatialFilters.jl:26-149).
//
// One pass: out = Σ_t w_t·val_t / Σ_t w_t·cnt_t over the 2w+1 taps t = -w..w, accumulated in tap order and finished
// with a true division, so the degenerate-width identity (σ << Δ: the result equals the input bit for bit,
// test/test_spatial_filters.jl:695-712) holds exactly.  Policies: Periodic wraps the index once, Shrink drops
// out-of-range taps from both sums, Edge clamps, Constant pads with (left, right).  Filters never read halos.
//
// Layout: threads run along x (coalesced) for every pass direction; the y and z passes walk rows / planes of
// the same x segment, each row read coalesced and re-used from L1/L2 by the neighbouring outputs of the block.
// Intermediates live in stream-ordered scratch owned by the library.
#include "ocb_common.cuh"
#include <math.h>
#include <stdlib.h>
#include <type_traits>

template <class T>
struct PassParams {
  const T* src; long long ss[3];      // source, pre-offset so that the 1-based index (i,j,k) is src[i*ss0 + j*ss1 + k*ss2]
  T* dst; long long ds[3];            // destination, same convention
  int lo[3], hi[3];                   // 1-based inclusive output window
  int dim, kind, width, policy, extent;
  int pairs_ok;                       // the source has an element on either side of every row (halo or padded scratch)
  T left, right, sigma, period;
  const T* nodes; const T* spacings;  // pre-offset: index m (1-based) -> nodes[m]
  T weights[2 * OCB_MAX_FILTER_WIDTH + 1];
};

template <class T> __device__ __forceinline__ T ocb_exp(T x);
template <> __device__ __forceinline__ double ocb_exp<double>(double x) { return exp(x); }
template <> __device__ __forceinline__ float ocb_exp<float>(float x) { return expf(x); }

// One 1-D pass, specialised at compile time on the filter kind and the boundary policy.  Outputs whose whole
// stencil lies inside the axis (all but 2w planes / rows / columns) take the branch-free path: 2w+1 loads and FMAs
// against weights staged in shared memory, divided by the weight sum accumulated in the same tap order.  Only the
// outputs next to a boundary run the per-tap policy logic.
template <class T, int KIND, int POLICY>
__global__ void __launch_bounds__(256) ocb_filter_pass_kernel(const __grid_constant__ PassParams<T> P) {
  __shared__ T wsh[2 * OCB_MAX_FILTER_WIDTH + 1];
  __shared__ T wtot_sh;
  const int n = P.extent, w = P.width, ntap = 2 * w + 1;
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  if (KIND == OCB_FILTER_GAUSSIAN) {
    if (tid < ntap) wsh[tid] = P.weights[tid];
    __syncthreads();
    if (tid == 0) { T a = T(0); for (int t = 0; t < ntap; ++t) a += wsh[t] * T(1); wtot_sh = a; }
    __syncthreads();
  }
  const T wtot = KIND == OCB_FILTER_GAUSSIAN ? wtot_sh : T(ntap);
  const int i = P.lo[0] + blockIdx.x * blockDim.x + threadIdx.x;
  const int j = P.lo[1] + blockIdx.y * blockDim.y + threadIdx.y;
  if (i > P.hi[0] || j > P.hi[1]) return;
  const int dim = P.dim;
  const long long sd = P.ss[dim];
  const T sig2 = T(2) * P.sigma * P.sigma;
  for (int k = P.lo[2] + blockIdx.z; k <= P.hi[2]; k += gridDim.z) {
    const int c = dim == 0 ? i : (dim == 1 ? j : k);                           // index along the filtered axis
    const T* base = P.src + i * P.ss[0] + j * P.ss[1] + k * P.ss[2] - c * sd;  // column along the filtered axis
    T out;
    if (KIND != OCB_FILTER_GAUSSIAN_STRETCHED && c - w >= 1 && c + w <= n) {
      const T* q = base + (long long)(c - w) * sd;
      T s = T(0);
      if (KIND == OCB_FILTER_BOX) {
        for (int t = 0; t < ntap; ++t, q += sd) s += __ldg(q);
      } else {
#pragma unroll 4
        for (int t = 0; t < ntap; ++t, q += sd) s += wsh[t] * __ldg(q);
      }
      out = s / wtot;
    } else {
      T s = T(0), wsum = T(0);
      int cnt_total = 0;
      T x0 = T(0);
      if (KIND == OCB_FILTER_GAUSSIAN_STRETCHED) x0 = __ldg(P.nodes + c);
      for (int t = -w; t <= w; ++t) {
        const int m = c + t;
        int cnt = 1;
        const bool below = m < 1, above = m > n;
        int mr;
        if (POLICY == OCB_POLICY_PERIODIC) mr = m + (below ? n : 0) - (above ? n : 0);
        else mr = below ? 1 : (above ? n : m);
        T val = __ldg(base + mr * sd);
        if (POLICY == OCB_POLICY_SHRINK) { if (below || above) { val = T(0); cnt = 0; } }
        else if (POLICY == OCB_POLICY_CONSTANT) { if (below) val = P.left; else if (above) val = P.right; }
        if (KIND == OCB_FILTER_BOX) {
          s += val;
          cnt_total += cnt;
        } else {
          T wt;
          if (KIND == OCB_FILTER_GAUSSIAN) {
            wt = wsh[t + w];
          } else {
            T x = __ldg(P.nodes + mr);
            if (POLICY == OCB_POLICY_PERIODIC) x = x - (below ? P.period : T(0)) + (above ? P.period : T(0));
            const T dxm = x - x0;
            wt = __ldg(P.spacings + mr) * ocb_exp<T>(-(dxm * dxm) / sig2);
          }
          s += wt * val;
          wsum += wt * T(cnt);
        }
      }
      out = (KIND == OCB_FILTER_BOX) ? s / T(cnt_total) : s / wsum;
    }
    P.dst[i * P.ds[0] + j * P.ds[1] + k * P.ds[2]] = out;
  }
}

// ---- interior fast paths (box and uniform Gaussian, half width W known at compile time) -------------------------
// Interior outputs (whole stencil inside the axis) need no boundary policy.  With W a template parameter the tap loops
// unroll completely and the weights are read as constant-bank operands straight from the kernel parameters.
//
// (1) x pass: one output per thread, 2W+1 cached loads of consecutive addresses.
template <class T, int W, bool BOX>
__global__ void __launch_bounds__(256) ocb_filter_x_kernel(const __grid_constant__ PassParams<T> P) {
  constexpr int NT = 2 * W + 1;
  const int i = P.lo[0] + blockIdx.x * blockDim.x + threadIdx.x;
  const int j = P.lo[1] + blockIdx.y * blockDim.y + threadIdx.y;
  if (i > P.hi[0] || j > P.hi[1]) return;
  T wtot = T(0);
  if (!BOX) {
#pragma unroll
    for (int t = 0; t < NT; ++t) wtot += P.weights[t] * T(1);
  }
  for (int k = P.lo[2] + blockIdx.z; k <= P.hi[2]; k += gridDim.z) {
    const T* q = P.src + (long long)(i - W) * P.ss[0] + j * P.ss[1] + k * P.ss[2];
    T s = T(0);
#pragma unroll
    for (int t = 0; t < NT; ++t) s = BOX ? s + __ldg(q + t) : s + P.weights[t] * __ldg(q + t);
    P.dst[i * P.ds[0] + j * P.ds[1] + k * P.ds[2]] = BOX ? s / T(NT) : s / wtot;
  }
}

// (2) y / z pass: a thread slides a register window of 2W+1 inputs along the filtered axis -- one load, 2W+1 FMAs and
// one store per output.  Threads run along x (coalesced); blockIdx.y walks the third axis, blockIdx.z the segments of
// the filtered axis.  The window rotates through compile-time slot indices (the loop is unrolled 2W+1 times).
// WRAP: the axis is periodic -- the kernel covers every output of the window and rows / planes beyond the ends are read at
// their wrapped index (no boundary-band launches for a periodic pass).
template <class T, int W, bool BOX, bool WRAP = false>
__global__ void __launch_bounds__(128) ocb_filter_slide_kernel(const __grid_constant__ PassParams<T> P, int c_lo, int c_hi, int seg) {
  constexpr int NT = 2 * W + 1;
  const int next = P.extent;
  auto wr = [&](int r) { return WRAP ? (r < 1 ? r + next : (r > next ? r - next : r)) : r; };
  const int dim = P.dim, oth = dim == 1 ? 2 : 1;                 // filtered axis, and the remaining (non-x) axis
  const int i = P.lo[0] + blockIdx.x * blockDim.x + threadIdx.x;
  const int o = P.lo[oth] + blockIdx.y;
  if (i > P.hi[0]) return;
  const int c0 = c_lo + blockIdx.z * seg;
  const int c1 = min(c0 + seg - 1, c_hi);
  if (c0 > c1) return;
  const long long sd = P.ss[dim], dd = P.ds[dim];
  const T* in = P.src + i * P.ss[0] + (long long)o * P.ss[oth];
  T* out = P.dst + i * P.ds[0] + (long long)o * P.ds[oth];
  T wtot = T(0);
  if (!BOX) {
#pragma unroll
    for (int t = 0; t < NT; ++t) wtot += P.weights[t] * T(1);
  }
  T win[NT];
#pragma unroll
  for (int m = 0; m < NT - 1; ++m) win[m] = __ldg(in + (long long)wr(c0 - W + m) * sd);
  win[NT - 1] = T(0);
  for (int c = c0; c <= c1; c += NT) {
    // the group's 2W+1 new inputs first, as independent loads: one load per output in program order leaves a warp with a
    // single request in flight, and the pass then runs at memory latency (0.35 ms) instead of bandwidth
    T nxt[NT];
#pragma unroll
    for (int t = 0; t < NT; ++t) nxt[t] = (c + t <= c1) ? __ldg(in + (long long)wr(c + t + W) * sd) : T(0);
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      if (c + t <= c1) {
        win[(2 * W + t) % NT] = nxt[t];
        T s = T(0);
#pragma unroll
        for (int j = 0; j < NT; ++j) s = BOX ? s + win[(t + j) % NT] : s + P.weights[j] * win[(t + j) % NT];
        out[(long long)(c + t) * dd] = BOX ? s / T(NT) : s / wtot;
      }
    }
  }
}

// (2b) x and y passes FUSED for periodic x and y (the common horizontal case, BASELINE config 5): a thread owns one x column and
// slides along y; for every new row it forms the x-filtered value of its column (2W+1 cached loads of consecutive addresses,
// the division included) and pushes it into the register window of the y pass.  The intermediate field of the staged form
// (16 B/cell of traffic and one launch) never exists.  Every output goes through exactly the staged arithmetic -- x sum in tap
// order, division, y sum in tap order, division -- so the results are bit-identical to the two staged passes.  Periodic wrap is
// native: rows wrap by index, columns whose stencil crosses the x boundary take per-tap wrapped loads (the edge warps only).
// EDGE = false: the columns whose x stencil lies inside the row (straight loads; the other threads leave at once);
// EDGE = true: the 2W columns next to the x boundary, per-tap wrapped loads (launched over those columns only).
template <class T, int W, bool BOX, bool EDGE>
__global__ void __launch_bounds__(128) ocb_filter_xy_kernel(const __grid_constant__ PassParams<T> P, int nx, int ny, int seg) {
  constexpr int NT = 2 * W + 1;
  int i = P.lo[0] + blockIdx.x * blockDim.x + threadIdx.x;
  if (EDGE) {                                               // thread e of the edge launch: columns 1..W, then nx-W+1..nx
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= 2 * W) return;
    i = e < W ? e + 1 : nx - 2 * W + e + 1;
    if (i < P.lo[0] || i > P.hi[0]) return;
  } else {
    if (i > P.hi[0] || i - W < 1 || i + W > nx) return;
  }
  const int k = P.lo[2] + blockIdx.y;
  const int c0 = P.lo[1] + blockIdx.z * seg;
  const int c1 = min(c0 + seg - 1, P.hi[1]);
  if (c0 > c1) return;
  T wtot = T(0);
  if (!BOX) {
#pragma unroll
    for (int t = 0; t < NT; ++t) wtot += P.weights[t] * T(1);
  }
  const T* plane = P.src + (long long)k * P.ss[2];
  T* out = P.dst + i * P.ds[0] + (long long)k * P.ds[2];
  // the x-filtered value of column i in row r (1-based, already wrapped into 1..ny)
  auto xrow = [&](int r) -> T {
    const T* row = plane + (long long)r * P.ss[1];
    T s = T(0);
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      int ii = i + t - W;
      if (EDGE) ii += ii < 1 ? nx : (ii > nx ? -nx : 0);
      s = BOX ? s + __ldg(row + ii) : s + P.weights[t] * __ldg(row + ii);
    }
    return BOX ? s / T(NT) : s / wtot;
  };
  auto wrap = [&](int r) { return r < 1 ? r + ny : (r > ny ? r - ny : r); };
  T win[NT];
#pragma unroll
  for (int m = 0; m < NT - 1; ++m) win[m] = xrow(wrap(c0 - W + m));
  win[NT - 1] = T(0);
  for (int c = c0; c <= c1; c += NT) {
    // the group's 2W+1 new x-filtered rows first: their loads are independent and overlap (one row at a time leaves a warp
    // with a single DRAM request in flight)
    T nxt[NT];
#pragma unroll
    for (int t = 0; t < NT; ++t) nxt[t] = (c + t <= c1) ? xrow(wrap(c + t + W)) : T(0);
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      if (c + t <= c1) {
        win[(2 * W + t) % NT] = nxt[t];
        T s = T(0);
#pragma unroll
        for (int j = 0; j < NT; ++j) s = BOX ? s + win[(t + j) % NT] : s + P.weights[j] * win[(t + j) % NT];
        out[(long long)(c + t) * P.ds[1]] = BOX ? s / T(NT) : s / wtot;
      }
    }
  }
}
template <class T, int W, bool BOX>
static void launch_xy_w(const PassParams<T>& P, int nx, int ny, cudaStream_t st) {
  const int ni = P.hi[0] - P.lo[0] + 1, nj = P.hi[1] - P.lo[1] + 1, nk = P.hi[2] - P.lo[2] + 1;
  int seg = 256;
  if (nj < seg) seg = nj;
  const int nseg = (nj + seg - 1) / seg;
  ocb_filter_xy_kernel<T, W, BOX, true><<<dim3(1, nk, nseg), dim3(32, 1, 1), 0, st>>>(P, nx, ny, seg);        // 2W <= 16 edge columns
  ocb_filter_xy_kernel<T, W, BOX, false><<<dim3((ni + 127) / 128, nk, nseg), dim3(128, 1, 1), 0, st>>>(P, nx, ny, seg);
}
template <class T, bool BOX>
static bool launch_xy(const PassParams<T>& P, int nx, int ny, cudaStream_t st) {
  switch (P.width) {
    case 1: launch_xy_w<T, 1, BOX>(P, nx, ny, st); return true;
    case 2: launch_xy_w<T, 2, BOX>(P, nx, ny, st); return true;
    case 3: launch_xy_w<T, 3, BOX>(P, nx, ny, st); return true;
    case 4: launch_xy_w<T, 4, BOX>(P, nx, ny, st); return true;
    case 5: launch_xy_w<T, 5, BOX>(P, nx, ny, st); return true;
    case 6: launch_xy_w<T, 6, BOX>(P, nx, ny, st); return true;
    case 8: launch_xy_w<T, 8, BOX>(P, nx, ny, st); return true;
    default: return false;
  }
}

// (3) Float64 pair kernel for the x pass: every thread owns two x-adjacent outputs and reads its 2W+2 inputs as W+1
// 16-byte pairs, which halves the L1 traffic the pass is bound by (9.7 -> 4.9 GB at config 5; 0.35 -> 0.28 ms).  The
// arithmetic per output is the one-column kernel's (same taps, same order, same division): results are bit-identical.
// Pairs are aligned in the SOURCE array (element index even relative to a 16-byte boundary; the caller checks that the
// base is 16-byte aligned and the y / z pitches are even); a pair that sticks out of the column window computes only
// its inside column.
__device__ __forceinline__ double2 ldg2(const double* p) { return __ldg(reinterpret_cast<const double2*>(p)); }

// x pass: outputs (i, i+1) need the 2W+2 inputs i-W .. i+1+W = W+1 aligned pairs.
// WRAP: x is periodic -- pairs whose window crosses an end of the row read their taps one by one at the wrapped column.
template <int W, bool BOX, bool WRAP = false>
__global__ void __launch_bounds__(256) ocb_filter_x2_kernel(const __grid_constant__ PassParams<double> P, int i_first) {
  constexpr int NT = 2 * W + 1;
  const int i = i_first + 2 * (blockIdx.x * blockDim.x + threadIdx.x);      // i - W is an aligned element
  const int j = P.lo[1] + blockIdx.y * blockDim.y + threadIdx.y;
  if (i > P.hi[0] || j > P.hi[1]) return;
  const bool v0 = i >= P.lo[0], v1 = i + 1 <= P.hi[0];
  const int nx = P.extent;
  const bool crossing = WRAP && (i - W < 1 || i + 1 + W > nx);
  double wtot = 0.0;
  if (!BOX) {
#pragma unroll
    for (int t = 0; t < NT; ++t) wtot += P.weights[t] * 1.0;
  }
  for (int k = P.lo[2] + blockIdx.z; k <= P.hi[2]; k += gridDim.z) {
    const double* q = P.src + (long long)(i - W) + j * P.ss[1] + k * P.ss[2];
    double v[2 * W + 2];
    if (crossing) {
      const double* row = P.src + j * P.ss[1] + k * P.ss[2];
#pragma unroll
      for (int m = 0; m < 2 * W + 2; ++m) { int ii = i - W + m; ii += ii < 1 ? nx : (ii > nx ? -nx : 0); v[m] = __ldg(row + ii); }
    } else {
#pragma unroll
      for (int m = 0; m <= W; ++m) { const double2 d = ldg2(q + 2 * m); v[2 * m] = d.x; v[2 * m + 1] = d.y; }
    }
    double s0 = 0.0, s1 = 0.0;
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      s0 = BOX ? s0 + v[t] : s0 + P.weights[t] * v[t];
      s1 = BOX ? s1 + v[t + 1] : s1 + P.weights[t] * v[t + 1];
    }
    double* o = P.dst + i * P.ds[0] + j * P.ds[1] + k * P.ds[2];
    if (v0) o[0] = BOX ? s0 / double(NT) : s0 / wtot;
    if (v1) o[P.ds[0]] = BOX ? s1 / double(NT) : s1 / wtot;
  }
}

// can the pair kernels read this pass's source?  16-byte aligned base, unit x stride, even y / z pitches
static bool pair_loads_ok(const PassParams<double>& P) {
  return P.pairs_ok && P.ss[0] == 1 && (P.ss[1] & 1) == 0 && (P.ss[2] & 1) == 0 && ((uintptr_t)P.src & 7) == 0 && !getenv("OCB_FILTER_NO_PAIRS");
}
template <class T> static bool pair_loads_ok_any(const PassParams<T>&) { return false; }
template <> bool pair_loads_ok_any<double>(const PassParams<double>& P) { return pair_loads_ok(P); }
// the first aligned element at or below index `i` of a row (rows share their parity because the pitches are even)
static int aligned_at_or_below(const PassParams<double>& P, int i) {
  const uintptr_t a = (uintptr_t)(P.src + i);
  return (a & 15) ? i - 1 : i;
}
template <int W, bool BOX>
static bool launch_pairs(const PassParams<double>& P, int c_lo, int c_hi, cudaStream_t st, bool wrap) {
  if (!pair_loads_ok(P)) return false;
  PassParams<double> Q = P;
  if (P.dim == 0) {
    Q.lo[0] = c_lo; Q.hi[0] = c_hi;
    const int i_first = aligned_at_or_below(P, c_lo - W) + W;               // i_first - W aligned, i_first <= c_lo
    const int npair = (c_hi - i_first) / 2 + 1, nj = P.hi[1] - P.lo[1] + 1, nk = P.hi[2] - P.lo[2] + 1;
    // the first pair may reach one element below c_lo - W and the last one above c_hi + W: both exist (c_lo - W >= 1 means the
    // row has an element at index 0 only as a halo or as the previous row's end; the value is loaded but never used)
    dim3 block(64, 4, 1), grid((npair + 63) / 64, (nj + 3) / 4, nk < 128 ? nk : 128);
    if (wrap) ocb_filter_x2_kernel<W, BOX, true><<<grid, block, 0, st>>>(Q, i_first);
    else ocb_filter_x2_kernel<W, BOX><<<grid, block, 0, st>>>(Q, i_first);
  } else {
    return false;                      // y / z passes: the one-column sliding-window kernel (a two-column window needs 108
                                       // registers and measured slower: 0.42 vs 0.35 ms)
  }
  return true;
}
template <class T, int W, bool BOX> struct PairLauncher { static bool go(const PassParams<T>&, int, int, cudaStream_t, bool) { return false; } };
template <int W, bool BOX> struct PairLauncher<double, W, BOX> {
  static bool go(const PassParams<double>& P, int c_lo, int c_hi, cudaStream_t st, bool wrap) { return launch_pairs<W, BOX>(P, c_lo, c_hi, st, wrap); }
};

// can the fast kernels of this pass cover a periodic axis end to end (wrapped reads instead of boundary bands)?
template <class T>
static bool wrap_capable(const PassParams<T>& P) {
  // measured at 512 x 512 x 256, 17 taps: 1.067 ms with the wrapped kernels against 0.967 ms with the interior kernels + band
  // launches on side streams (the wrap arithmetic in every load of the sliding window costs more than the bands): opt-in only
  if (P.policy != OCB_POLICY_PERIODIC || !getenv("OCB_FILTER_WRAP")) return false;
  if (P.dim == 0) return std::is_same<T, double>::value && pair_loads_ok_any(P);
  return true;
}

template <class T, int W, bool BOX>
static void launch_fast_w(const PassParams<T>& P, int c_lo, int c_hi, cudaStream_t st, bool wrap) {
  if (PairLauncher<T, W, BOX>::go(P, c_lo, c_hi, st, wrap)) return;
  PassParams<T> Q = P;
  if (P.dim == 0) {
    Q.lo[0] = c_lo; Q.hi[0] = c_hi;
    const int ni = c_hi - c_lo + 1, nj = P.hi[1] - P.lo[1] + 1, nk = P.hi[2] - P.lo[2] + 1;
    dim3 block(64, 4, 1), grid((ni + 63) / 64, (nj + 3) / 4, nk < 128 ? nk : 128);
    ocb_filter_x_kernel<T, W, BOX><<<grid, block, 0, st>>>(Q);
  } else {
    const int oth = P.dim == 1 ? 2 : 1;
    const int ni = P.hi[0] - P.lo[0] + 1, no = P.hi[oth] - P.lo[oth] + 1, nc = c_hi - c_lo + 1;
    int seg = 128;
    if (nc < seg) seg = nc;
    dim3 block(128, 1, 1), grid((ni + 127) / 128, no, (nc + seg - 1) / seg);
    if (wrap) ocb_filter_slide_kernel<T, W, BOX, true><<<grid, block, 0, st>>>(Q, c_lo, c_hi, seg);
    else ocb_filter_slide_kernel<T, W, BOX><<<grid, block, 0, st>>>(Q, c_lo, c_hi, seg);
  }
}
static bool fast_width_compiled(int w) { return (w >= 1 && w <= 6) || w == 8; }
// returns false when no fast instantiation exists for this half width
template <class T, bool BOX>
static bool launch_fast(const PassParams<T>& P, int c_lo, int c_hi, cudaStream_t st, bool wrap = false) {
  switch (P.width) {
    case 1: launch_fast_w<T, 1, BOX>(P, c_lo, c_hi, st, wrap); return true;
    case 2: launch_fast_w<T, 2, BOX>(P, c_lo, c_hi, st, wrap); return true;
    case 3: launch_fast_w<T, 3, BOX>(P, c_lo, c_hi, st, wrap); return true;
    case 4: launch_fast_w<T, 4, BOX>(P, c_lo, c_hi, st, wrap); return true;
    case 5: launch_fast_w<T, 5, BOX>(P, c_lo, c_hi, st, wrap); return true;
    case 6: launch_fast_w<T, 6, BOX>(P, c_lo, c_hi, st, wrap); return true;
    case 8: launch_fast_w<T, 8, BOX>(P, c_lo, c_hi, st, wrap); return true;
    default: return false;
  }
}

template <class T, int KIND>
static void launch_pass_policy(const PassParams<T>& P, dim3 grid, dim3 block, cudaStream_t st) {
  switch (P.policy) {
    case OCB_POLICY_PERIODIC: ocb_filter_pass_kernel<T, KIND, OCB_POLICY_PERIODIC><<<grid, block, 0, st>>>(P); break;
    case OCB_POLICY_SHRINK: ocb_filter_pass_kernel<T, KIND, OCB_POLICY_SHRINK><<<grid, block, 0, st>>>(P); break;
    case OCB_POLICY_EDGE: ocb_filter_pass_kernel<T, KIND, OCB_POLICY_EDGE><<<grid, block, 0, st>>>(P); break;
    default: ocb_filter_pass_kernel<T, KIND, OCB_POLICY_CONSTANT><<<grid, block, 0, st>>>(P); break;
  }
}
template <class T>
static void launch_pass(const PassParams<T>& P, dim3 grid, dim3 block, cudaStream_t st) {
  if (P.kind == OCB_FILTER_BOX) launch_pass_policy<T, OCB_FILTER_BOX>(P, grid, block, st);
  else if (P.kind == OCB_FILTER_GAUSSIAN) launch_pass_policy<T, OCB_FILTER_GAUSSIAN>(P, grid, block, st);
  else launch_pass_policy<T, OCB_FILTER_GAUSSIAN_STRETCHED>(P, grid, block, st);
}

// two library-owned side streams (per host thread and device) for the boundary bands of a pass
struct BandStreams { int device; cudaStream_t s[2]; cudaEvent_t fork, join[2]; };
static BandStreams* band_streams() {
  static thread_local BandStreams bs = {-1, {nullptr, nullptr}, nullptr, {nullptr, nullptr}};
  if (getenv("OCB_FILTER_NO_SIDE_STREAMS")) return nullptr;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
  if (bs.device == dev) return &bs;
  if (bs.device >= 0) return nullptr;                   // this thread already serves another device: stay on one stream
  bool ok = cudaStreamCreateWithFlags(&bs.s[0], cudaStreamNonBlocking) == cudaSuccess &&
            cudaStreamCreateWithFlags(&bs.s[1], cudaStreamNonBlocking) == cudaSuccess &&
            cudaEventCreateWithFlags(&bs.fork, cudaEventDisableTiming) == cudaSuccess &&
            cudaEventCreateWithFlags(&bs.join[0], cudaEventDisableTiming) == cudaSuccess &&
            cudaEventCreateWithFlags(&bs.join[1], cudaEventDisableTiming) == cudaSuccess;
  if (!ok) { cudaGetLastError(); return nullptr; }
  bs.device = dev;
  return &bs;
}

template <class T>
static int filter_impl(const ocb_field* src, const int32_t ext[3], const ocb_filter_pass* passes, int n_pass,
                       const ocb_field* dst, const int32_t wlo[3], const int32_t whi[3], cudaStream_t st) {
  // window of the final pass
  int lo[3], hi[3];
  for (int d = 0; d < 3; ++d) {
    lo[d] = (wlo && (wlo[d] || whi[d])) ? wlo[d] : 1;
    hi[d] = (whi && (wlo[d] || whi[d])) ? whi[d] : ext[d];
    OCB_REQUIRE(lo[d] >= 1 && hi[d] <= ext[d] && lo[d] <= hi[d], "filter window [%d, %d] outside 1..%d along axis %d", lo[d], hi[d], ext[d], d + 1);
  }
  const size_t n_full = (size_t)ext[0] * ext[1] * ext[2];
  T* tmp[2] = {nullptr, nullptr};
  // two spare elements in front of and behind each temporary: the pair kernels read whole 16-byte pairs
  T* tmp_alloc[2] = {nullptr, nullptr};
  for (int t = 0; t < n_pass - 1 && t < 2; ++t) {
    OCB_CUDA(cudaMallocAsync((void**)&tmp_alloc[t], (n_full + 4) * sizeof(T), st));
    tmp[t] = tmp_alloc[t] + 2;
  }
  int rc = OCB_OK;
  // x and y passes of the same periodic box / uniform Gaussian filter run as ONE kernel (ocb_filter_xy_kernel)
  bool fuse_xy = n_pass >= 2 && passes[0].dim == 1 && passes[1].dim == 2 && passes[0].kind == passes[1].kind &&
                 passes[0].kind != OCB_FILTER_GAUSSIAN_STRETCHED && passes[0].policy == OCB_POLICY_PERIODIC &&
                 passes[1].policy == OCB_POLICY_PERIODIC && passes[0].width == passes[1].width && fast_width_compiled(passes[0].width) &&
                 ext[0] >= 2 * passes[0].width + 1 && ext[1] >= 2 * passes[0].width + 1 && src->stride[0] == 1 && getenv("OCB_FILTER_XY");
  // (measured at 512 x 512 x 256, 17 taps: the fused kernel 0.93 + 0.28 ms against 0.28 + 0.24 ms for the two staged passes --
  //  2W+1 cached loads AND a 2W+1-deep register window per output leave too few loads in flight; kept behind OCB_FILTER_XY=1,
  //  bit-identical, for tuning runs: profiles/r2_summary.md)
  if (fuse_xy && passes[0].kind == OCB_FILTER_GAUSSIAN)
    for (int t = 0; t < 2 * passes[0].width + 1; ++t) if (T(passes[0].weights[t]) != T(passes[1].weights[t])) fuse_xy = false;
  for (int ps = 0; ps < n_pass; ++ps) {
    const ocb_filter_pass& fp = passes[ps];
    PassParams<T> P;
    memset(&P, 0, sizeof(P));
    const bool fused_now = fuse_xy && ps == 0;
    const bool first = ps == 0;
    if (fused_now) ps = 1;                                 // this iteration does the work of passes 0 and 1
    const bool last = ps == n_pass - 1;
    if (first) {
      long long off = 0;
      for (int d = 0; d < 3; ++d) { P.ss[d] = src->stride[d]; off += (long long)(src->halo[d] - 1) * src->stride[d]; }
      P.src = (const T*)src->data + off;
      P.pairs_ok = src->halo[0] >= 1 && src->halo[1] >= 1 && src->halo[2] >= 1;
    } else {                       // dense scratch indexed from 1
      P.pairs_ok = 1;
      P.ss[0] = 1; P.ss[1] = ext[0]; P.ss[2] = (long long)ext[0] * ext[1];
      P.src = tmp[(ps - 1) & 1] - (P.ss[0] + P.ss[1] + P.ss[2]);
    }
    if (last) {
      long long off = 0;
      for (int d = 0; d < 3; ++d) { P.ds[d] = dst->stride[d]; off += ((long long)dst->halo[d] - lo[d]) * dst->stride[d]; P.lo[d] = lo[d]; P.hi[d] = hi[d]; }
      P.dst = (T*)dst->data + off;
    } else {
      P.ds[0] = 1; P.ds[1] = ext[0]; P.ds[2] = (long long)ext[0] * ext[1];
      P.dst = tmp[ps & 1] - (P.ds[0] + P.ds[1] + P.ds[2]);
      for (int d = 0; d < 3; ++d) { P.lo[d] = 1; P.hi[d] = ext[d]; }
    }
    P.dim = fp.dim - 1; P.kind = fp.kind; P.width = fp.width; P.policy = fp.policy; P.extent = fp.extent;
    P.left = T(fp.left); P.right = T(fp.right); P.sigma = T(fp.sigma); P.period = T(fp.period);
    if (fp.kind == OCB_FILTER_GAUSSIAN) for (int t = 0; t < 2 * fp.width + 1; ++t) P.weights[t] = T(fp.weights[t]);
    if (fused_now) {
      const bool done = fp.kind == OCB_FILTER_BOX ? launch_xy<T, true>(P, ext[0], ext[1], st) : launch_xy<T, false>(P, ext[0], ext[1], st);
      if (!done || cudaGetLastError() != cudaSuccess) { ocb_set_error("fused x-y filter launch failed"); rc = OCB_ERR_CUDA; break; }
      continue;
    }
    if (fp.kind == OCB_FILTER_GAUSSIAN_STRETCHED) { P.nodes = (const T*)fp.nodes - 1; P.spacings = (const T*)fp.spacings - 1; }
    auto launch_generic = [&](const PassParams<T>& Q, cudaStream_t on) {
      const int ni = Q.hi[0] - Q.lo[0] + 1, nj = Q.hi[1] - Q.lo[1] + 1, nk = Q.hi[2] - Q.lo[2] + 1;
      if (ni <= 0 || nj <= 0 || nk <= 0) return;
      int bx = 32;                                        // a band a few columns wide: narrower blocks, no idle lanes
      while (bx > 4 && bx / 2 >= ni) bx /= 2;
      const int by = 256 / bx;
      dim3 block(bx, by, 1), grid((ni + bx - 1) / bx, (nj + by - 1) / by, nk < 256 ? nk : 256);
      launch_pass<T>(Q, grid, block, on);
    };
    // interior outputs [w+1, n-w] of the window go to the unrolled fast kernels, the two boundary bands to the
    // policy-aware kernel
    const int dm = P.dim;
    const int c_lo = P.lo[dm] > fp.width + 1 ? P.lo[dm] : fp.width + 1;
    const int c_hi = P.hi[dm] < fp.extent - fp.width ? P.hi[dm] : fp.extent - fp.width;
    bool fast = false;
    BandStreams* bs = nullptr;
    if (fp.kind != OCB_FILTER_GAUSSIAN_STRETCHED && c_hi - c_lo + 1 >= 4 * fp.width + 2 && !getenv("OCB_FILTER_NO_FAST")) {
      // the two boundary bands write other outputs than the interior kernel: they run next to it on two side streams
      // (fork here, after the previous pass; join below, before the next pass)
      bs = band_streams();
      if (bs) {
        cudaEventRecord(bs->fork, st);
        cudaStreamWaitEvent(bs->s[0], bs->fork, 0);
        cudaStreamWaitEvent(bs->s[1], bs->fork, 0);
      }
      fast = fast_width_compiled(fp.width);
    }
    if (fast && wrap_capable<T>(P) && P.lo[dm] == 1 && P.hi[dm] == fp.extent && fp.extent >= 2 * fp.width + 1) {
      // a periodic axis covered end to end: the fast kernel reads beyond the ends at the wrapped index, no boundary bands
      if (fp.kind == OCB_FILTER_BOX) launch_fast<T, true>(P, P.lo[dm], P.hi[dm], st, true); else launch_fast<T, false>(P, P.lo[dm], P.hi[dm], st, true);
    } else if (fast) {
      // bands first: their few blocks start at once and the interior kernel fills the rest of the machine
      PassParams<T> B = P;
      B.hi[dm] = c_lo - 1; launch_generic(B, bs ? bs->s[0] : st);           // low band
      B = P; B.lo[dm] = c_hi + 1; launch_generic(B, bs ? bs->s[1] : st);    // high band
      if (fp.kind == OCB_FILTER_BOX) launch_fast<T, true>(P, c_lo, c_hi, st); else launch_fast<T, false>(P, c_lo, c_hi, st);
      if (bs) {
        for (int q = 0; q < 2; ++q) { cudaEventRecord(bs->join[q], bs->s[q]); cudaStreamWaitEvent(st, bs->join[q], 0); }
      }
    } else {
      launch_generic(P, st);
    }
    if (cudaGetLastError() != cudaSuccess) { ocb_set_error("filter pass launch failed"); rc = OCB_ERR_CUDA; break; }
  }
  for (int t = 0; t < 2; ++t) if (tmp_alloc[t]) cudaFreeAsync(tmp_alloc[t], st);
  return rc;
}

extern "C" int ocb_filter_execute(int32_t dtype, const ocb_field* src, const int32_t src_extent[3], const ocb_filter_pass* passes,
                                  int32_t n_pass, const ocb_field* dst, const int32_t window_lo[3], const int32_t window_hi[3], void* stream) {
  OCB_REQUIRE(src && dst && passes && src_extent, "ocb_filter_execute: null argument");
  OCB_REQUIRE(src->kind == OCB_FIELD_ARRAY && src->data && dst->kind == OCB_FIELD_ARRAY && dst->data, "filter operands must be array fields");
  OCB_REQUIRE(n_pass >= 1 && n_pass <= 3, "a staged filter has 1..3 passes; got %d", n_pass);
  OCB_REQUIRE(dtype == OCB_F32 || dtype == OCB_F64, "bad dtype");
  for (int p = 0; p < n_pass; ++p) {
    const ocb_filter_pass& fp = passes[p];
    OCB_REQUIRE(fp.dim >= 1 && fp.dim <= 3, "pass %d: dim must be 1, 2 or 3", p);
    for (int q = 0; q < p; ++q) OCB_REQUIRE(passes[q].dim != fp.dim, "passes %d and %d filter the same dim", q, p);
    OCB_REQUIRE(fp.width >= 1 && fp.width <= OCB_MAX_FILTER_WIDTH, "pass %d: half width %d outside 1..%d", p, fp.width, OCB_MAX_FILTER_WIDTH);
    OCB_REQUIRE(fp.kind >= OCB_FILTER_BOX && fp.kind <= OCB_FILTER_GAUSSIAN_STRETCHED, "pass %d: unknown filter kind", p);
    OCB_REQUIRE(fp.policy >= OCB_POLICY_PERIODIC && fp.policy <= OCB_POLICY_CONSTANT, "pass %d: unknown boundary policy", p);
    OCB_REQUIRE(fp.extent == src_extent[fp.dim - 1], "pass %d: extent %d differs from the operand's length %d", p, fp.extent, src_extent[fp.dim - 1]);
    if (fp.policy == OCB_POLICY_PERIODIC)   // SpatialFilters.jl:248-257 (a tap wraps at most once)
      OCB_REQUIRE(fp.width <= fp.extent, "for the periodic direction d=%d, `N` (%d) exceeds the maximum allowed 2*Nd_grid+1 = %d", fp.dim,
                  2 * fp.width + 1, 2 * fp.extent + 1);
    if (fp.kind == OCB_FILTER_GAUSSIAN_STRETCHED) OCB_REQUIRE(fp.nodes && fp.spacings && fp.sigma > 0, "pass %d: stretched Gaussian needs nodes, spacings and sigma", p);
  }
  int rc = ocb_require_device();
  if (rc != OCB_OK) return rc;
  OcbDeviceGuard device_guard(src->data);
  ocb_keep_scratch_pool();
  return dtype == OCB_F64 ? filter_impl<double>(src, src_extent, passes, n_pass, dst, window_lo, window_hi, (cudaStream_t)stream)
                          : filter_impl<float>(src, src_extent, passes, n_pass, dst, window_lo, window_hi, (cudaStream_t)stream);
}
