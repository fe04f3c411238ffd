// K3: sort-based reference height z✶ for the background potential energy
// (reference: compute!(::SortedReferenceHeightField), src/BackgroundPotentialEnergyEquation.jl:610-619, with
// rank_by_buoyancy! :310-316, slot_centers! :326-331, slot_faces :339-348, reference_potential_function :359-368,
// fill_reference_potential! :383-397, assign_reference_height! for ThreeDimensionalSort :413-426 and
// HeavisideIntegral :428-466).
//
// The reference calls sortperm! (Base on the CPU, CUDA.jl's on the GPU).  Here the sort is a hand-written stable LSD
// radix sort (8-bit digits) of order-preserving integer keys with a 32-bit index payload ("onesweep"):
//   once       histogram of all digits (the first pass's reader makes the keys from the caller's values / the field's interior)
//   per digit  ONE persistent kernel: ticketed tiles, ranking through per-warp mask words in shared memory, a two-level chained
//              scan over the tiles' digit counts (groups of 16 tickets, RED-accumulated group totals), software-pipelined so that a
//              tile's prefix is asked for one tile later; the last pass writes values and permutation into the caller's arrays
// (n >= 2^30 or OCB_SORT_TABLE_SCAN: per-pass histogram table + scan + scatter instead of the chained scan.)
// Stable means ties keep their original (column-major) order, which is what Base.sortperm! does on the CPU; the
// reference's GPU sort gives no such guarantee, hence only permutation-invariant results are compared.
// Key transform: flip all bits of negative floats, set the sign bit of non-negative ones -- Julia's isless order on
// finite values and on ±0 (-0.0 < +0.0).
//
// The epilogues (cumulative volume -> slot centres and faces, Ψ by a second scan and a binary search per cell,
// run detection with max / reversed-min scans for HeavisideIntegral, scatter back through the permutation) are
// block-scan kernels over the sorted arrays.  All scratch is stream-ordered and owned by the library.
#include "ocb_common.cuh"
#include <float.h>

namespace {

constexpr int RS_THREADS = 256;
constexpr int RS_ITEMS = 16;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;   // keys per block
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_MAX_DEVICES = 64;

template <class K> struct KeyTraits;
template <> struct KeyTraits<uint64_t> { using F = double; static constexpr int BITS = 64; };
template <> struct KeyTraits<uint32_t> { using F = float; static constexpr int BITS = 32; };

__device__ __forceinline__ uint64_t to_key(double v) {
  uint64_t u = (uint64_t)__double_as_longlong(v);
  return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double from_key(uint64_t k) {
  uint64_t u = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
  return __longlong_as_double((long long)u);
}
__device__ __forceinline__ uint32_t to_key(float v) {
  uint32_t u = __float_as_uint(v);
  return (u >> 31) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float from_key(uint32_t k) {
  uint32_t u = (k >> 31) ? (k & 0x7fffffffu) : ~k;
  return __uint_as_float(u);
}

template <class K>
__global__ void rs_unmake_keys(const K* __restrict__ keys, typename KeyTraits<K>::F* __restrict__ v, long long n) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t < n) v[t] = from_key(keys[t]);
}

// Where a pass reads its keys from.  The first pass of a sort reads the caller's data and makes the keys on the fly (and its
// payload is the key's position), so no make-keys pass writes and re-reads 12 B/key.  A thread walks its keys with a cursor
// (`at(t)`, then `advance(step)` between reads); `first` is the read for the digits, `again` the second read of the same key
// when the tile is staged (from L2).
template <class K> struct PtrKeys {
  const K* p;
  struct Cursor {
    const K* q;
    __device__ __forceinline__ K first() const { return __ldg(q); }
    __device__ __forceinline__ K again() const { return __ldcg(q); }
    __device__ __forceinline__ void advance(uint32_t step) { q += step; }
  };
  __device__ __forceinline__ Cursor at(long long t) const { return Cursor{p + t}; }
};
template <class K> struct FloatKeys {           // a flat floating-point array
  const typename KeyTraits<K>::F* p;
  struct Cursor {
    const typename KeyTraits<K>::F* q;
    __device__ __forceinline__ K first() const { return to_key(__ldg(q)); }
    __device__ __forceinline__ K again() const { return to_key(__ldcg(q)); }
    __device__ __forceinline__ void advance(uint32_t step) { q += step; }
  };
  __device__ __forceinline__ Cursor at(long long t) const { return Cursor{p + t}; }
};
// unsigned division by a run-time invariant (Granlund & Montgomery): q = (t1 + ((n - t1) >> sh1)) >> sh2 with t1 = mulhi(m, n)
struct FastDiv {
  uint32_t d, m, sh1, sh2;
  FastDiv() : d(1), m(1), sh1(0), sh2(0) {}
  explicit FastDiv(uint32_t div) : d(div) {
    uint32_t s = 0;
    while ((1ull << s) < div) ++s;
    m = (uint32_t)(((1ull << 32) * ((1ull << s) - div)) / div + 1);
    sh1 = s < 1 ? s : 1; sh2 = s - sh1;
  }
  __device__ __forceinline__ uint32_t div(uint32_t n) const { const uint32_t t1 = __umulhi(m, n); return (t1 + ((n - t1) >> sh1)) >> sh2; }
};
template <class T, class K> struct FieldKeys {  // the interior of a field (< 2^32 cells): b1 pre-offset to index (1,1,1) -> element 0
  const T* b1; long long s0, s1, s2; FastDiv dx, dy;
  struct Cursor {                               // (i, j) of the cell and its address; a step carries into j and k
    const T* q; uint32_t i, j, Nx, Ny; long long s0, back_x, back_y;
    __device__ __forceinline__ K first() const { return to_key(__ldg(q)); }
    __device__ __forceinline__ K again() const { return to_key(__ldcg(q)); }
    __device__ __forceinline__ void advance(uint32_t step) {
      i += step; q += step * s0;
      while (i >= Nx) {
        i -= Nx; q += back_x; ++j;              // back_x = s1 - Nx s0: to the start of the next row
        if (j == Ny) { j = 0; q += back_y; }    // back_y = s2 - Ny s1: to the start of the next plane
      }
    }
  };
  __device__ __forceinline__ Cursor at(long long t) const {
    const uint32_t tt = (uint32_t)t, q = dx.div(tt), i = tt - q * dx.d, k = dy.div(q), j = q - k * dy.d;
    return Cursor{b1 + (i * s0 + j * s1 + k * s2), i, j, dx.d, dy.d, s0, s1 - (long long)dx.d * s0, s2 - (long long)dy.d * s1};
  }
};
// ... and the value such a pattern stands for (the epilogues of the uniform-volume reference height read the sorted VALUES)
__device__ __forceinline__ double value_of_bits(uint64_t u) { return __longlong_as_double((long long)u); }
__device__ __forceinline__ float value_of_bits(uint32_t u) { return __uint_as_float(u); }
// the bit pattern of from_key(k): what the last pass of a sort stores when the caller wants values, not keys
__device__ __forceinline__ uint64_t unkey_bits(uint64_t k) { return (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k; }
__device__ __forceinline__ uint32_t unkey_bits(uint32_t k) { return (k >> 31) ? (k & 0x7fffffffu) : ~k; }
template <class K, class In>
__global__ void rs_materialize_keys(const In in, K* __restrict__ keys, uint32_t* __restrict__ idx, long long n) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t < n) { keys[t] = in.at(t).first(); idx[t] = (uint32_t)t; }
}

// per-tile digit histogram, written digit-major: table[digit * nblocks + block]
template <class K>
__global__ void __launch_bounds__(RS_THREADS) rs_histogram(const K* __restrict__ keys, long long n, int shift, uint32_t* __restrict__ table, int nblocks) {
  __shared__ uint32_t h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  const long long base = (long long)blockIdx.x * RS_TILE;
#pragma unroll
  for (int r = 0; r < RS_ITEMS; ++r) {
    const long long t = base + r * RS_THREADS + threadIdx.x;
    if (t < n) atomicAdd(&h[(uint32_t)(keys[t] >> shift) & 255u], 1u);
  }
  __syncthreads();
  table[(size_t)threadIdx.x * nblocks + blockIdx.x] = h[threadIdx.x];
}

// stable scatter of a tile: warp w owns the contiguous sub-tile [w*512, (w+1)*512), walked in 16 rounds of 32.
// The tile is first put in digit order in shared memory (keys in `stage`, the payload in `ibuf`), so that the global writes are
// runs of consecutive addresses per digit (16 keys = 128 B on average) instead of one sector per key.
// CHAINED = false: the tile is blockIdx.x and its global digit offsets come from the scanned per-tile histogram table (`offsets`,
// digit-major).
// CHAINED = true ("onesweep"): no histogram pass and no table scan -- `offsets` holds the 256 global digit bases of this pass
// (from ONE up-front histogram of all digits); the blocks are persistent, take tile numbers from a ticket counter, publish their
// digit counts and find their exclusive prefix from the counts of the tiles with smaller tickets (a tile only ever waits for
// those, and they are running).  A wait is bounded: a stuck wait sets desc's error word and traps (RS_SPIN_LIMIT) instead of
// hanging the device or returning a wrong order.
//
// Two-level chained scan.  Tiles form groups of RS_G consecutive tickets.  Every tile publishes its digit counts agg[tile][256]
// and adds them to its group's total gt[g][256] (RED.ADD), then bumps the group's counter; the block whose bump completes group g
// computes GX[g+1] = the counts of all tiles of groups <= g from the nearest published GX[g+1-k] and the totals of the complete
// groups in between (a decoupled look-back over GROUPS: 1.8 of them per microsecond instead of 28 tiles) and publishes it.  The
// exclusive prefix of tile t = (g, i) is GX[g] + the counts of the i tiles before it in its group: one poll of 16 flags and one
// batch of <= 16 independent loads.  A one-level look-back cannot keep up here: at 28 tiles per microsecond its inclusive
// frontier trails ~64 tiles behind (measured: 64 words per thread and 5 polls per tile).
//
// Software pipeline.  What a look-back costs is mostly WAITING: the tiles in flight progress at different speeds, and a tile that
// asks for its predecessors' counts right after publishing its own waits for the slowest of them (measured: 5 us of a 20 us tile,
// whatever the look-back's shape).  So the block publishes tile t's counts, keeps the staged tile in shared memory, loads and
// ranks tile t+1 (~7 us), and only then asks for tile t's prefix and writes it out: everything the prefix needs was published
// by the other blocks one tile earlier in THEIR loops.  Between the two the registers hold of tile t+1 only the digits (8 bits)
// and the in-warp ranks (16 bits) of the thread's 16 keys -- 12 registers; the keys are read again (from L2) when they are
// staged, together with the payload (prefetched into L2 at the start of the tile).
// desc layout (32-bit words): aggf[nblocks] flags, the ticket counter, the error word, gcount[ngroups], gxf[ngroups], padding to
// rs_desc_header(nblocks), gt[ngroups][256] -- all of that cleared before every pass (rs_desc_clear words) -- then
// agg[nblocks][256] and gx[ngroups][256], write-once per pass.
// A wait is bounded: ~2.7 s of polling (every tile waited for holds a smaller ticket and is running, so this only fires on a
// device fault) ends the kernel with a trap -- a sticky launch failure at the caller's next synchronisation, not a wrong sort.
constexpr unsigned RS_SPIN_LIMIT = 1u << 26;
constexpr int RS_G = 16;    // tiles per group
constexpr int RS_GK = 8;    // groups looked back over per poll
__host__ __device__ constexpr size_t rs_groups(int nblocks) { return ((size_t)nblocks + RS_G - 1) / RS_G; }
__host__ __device__ constexpr size_t rs_desc_header(int nblocks) { return (((size_t)nblocks + 2 + 2 * rs_groups(nblocks) + 63) / 64) * 64; }
__host__ __device__ constexpr size_t rs_desc_clear(int nblocks) { return rs_desc_header(nblocks) + 256ull * rs_groups(nblocks); }
constexpr size_t rs_table_words(int nblocks) {                         // also >= the 256 nblocks table of the scan path
  return rs_desc_header(nblocks) + 256ull * ((size_t)nblocks + 2 * rs_groups(nblocks));
}
__device__ __forceinline__ void rs_st_release(uint32_t* p, uint32_t v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// flag poll: a relaxed gpu-scope load (L2).  ld.acquire.gpu compiles to the same load plus CCTL.IVALL -- an invalidation of the
// whole L1 per poll and warp.  The words read after a flag are read with __ldcg (L2, never L1) at addresses / under predicates
// that depend on the flag's value, so they are issued after the flag has arrived, and the writer orders its words before its
// flag with a gpu-scope release.
__device__ __forceinline__ uint32_t rs_ld_flag(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// An optional second payload: one double per element, formed from the element's position when its tile is staged and moved with
// the pair (the bucketing pass of the reference-height epilogue carries the cell's Psi this way instead of gathering it later).
struct NoPay2 {
  static constexpr bool ON = false;
  __device__ __forceinline__ double at(long long, uint32_t) const { return 0.0; }
};
template <class K, class Pay2 = NoPay2> constexpr size_t rs_scatter_smem() {
  return sizeof(K) * RS_TILE + sizeof(uint32_t) * (RS_TILE + 2 * RS_WARPS * 256 + 256 + RS_WARPS + 8) + (Pay2::ON ? sizeof(double) * RS_TILE : 0);
}
template <class K, bool CHAINED = false, class In = PtrKeys<K>, bool UNMAKE = false, class Pay2 = NoPay2>
__global__ void __launch_bounds__(RS_THREADS, Pay2::ON ? 2 : 3) rs_scatter(const In keys_in, const uint32_t* __restrict__ idx_in, K* __restrict__ keys_out,
                                                          uint32_t* __restrict__ idx_out, long long n, int shift,
                                                          const uint32_t* __restrict__ offsets, int nblocks, uint32_t* desc = nullptr,
                                                          const Pay2 pay2 = Pay2(), double* __restrict__ pay2_out = nullptr) {
  extern __shared__ __align__(16) unsigned char rs_smem[];
  K* stage = reinterpret_cast<K*>(rs_smem);                                     // the staged tile's keys in digit order
  uint32_t* ibuf = reinterpret_cast<uint32_t*>(stage + RS_TILE);                // the staged tile's payload in digit order
  // per warp and digit: ranking mask | running count << 32, then (count half) the tile position of the warp's first key of the digit
  unsigned long long (*wmc)[256] = reinterpret_cast<unsigned long long (*)[256]>(ibuf + RS_TILE);
  uint32_t* gbase = reinterpret_cast<uint32_t*>(&wmc[RS_WARPS][0]);   // global position of the staged tile's first key of each digit, minus its tile position
  uint32_t* wtot = gbase + 256;           // [RS_WARPS] block-scan totals, [RS_WARPS] the next tile number
  double* ibuf2 = reinterpret_cast<double*>(wtot + RS_WARPS + 8);             // the staged tile's second payload (Pay2::ON)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1u;
  const int d = threadIdx.x;
  const int ngroups = (int)rs_groups(nblocks);
  uint32_t* const gcount = desc + nblocks + 2;
  uint32_t* const gxf = gcount + ngroups;
  uint32_t* const gt = desc + rs_desc_header(nblocks);
  uint32_t* const agg = gt + (size_t)ngroups * 256;
  uint32_t* const gx = agg + (size_t)nblocks * 256;
  const int wl = warp * (RS_ITEMS * 32) + lane;         // tile position of the thread's key of round r: wl + 32 r
  int tile = blockIdx.x;
  if (CHAINED) {
    if (threadIdx.x == 0) wtot[RS_WARPS] = atomicAdd(desc + nblocks, 1u);
    __syncthreads();
    tile = (int)wtot[RS_WARPS];
  }
  int ptile = -1;                   // the staged tile, waiting for its prefix
  uint32_t prun = 0, pstart = 0;    // its count and tile position of digit d
  for (;;) {
    const bool have = tile < nblocks;
    const long long tbase = (long long)tile * RS_TILE;
    const int ntile = have ? (int)((n - tbase) < (long long)RS_TILE ? (n - tbase) : (long long)RS_TILE) : 0;
    uint32_t dg[RS_ITEMS / 4], rk[RS_ITEMS / 2];        // the thread's 16 digits (8 bits each) and in-warp ranks (16 bits each)
    uint32_t run = 0, start = 0, closes = 0;
    if (have) {
      for (int q = lane; q < 256; q += 32) wmc[warp][q] = 0ull;
      if (idx_in && threadIdx.x * 32 < ntile) asm volatile("prefetch.global.L2 [%0];" ::"l"(idx_in + tbase + threadIdx.x * 32));
      {
        K key[RS_ITEMS];
        auto cur = keys_in.at(tbase + wl);
#pragma unroll
        for (int r = 0; r < RS_ITEMS; ++r) { key[r] = (wl + 32 * r) < ntile ? cur.first() : (K)~(K)0; cur.advance(32); }
#pragma unroll
        for (int q = 0; q < RS_ITEMS / 4; ++q) {
          dg[q] = 0;
#pragma unroll
          for (int b = 0; b < 4; ++b) dg[q] |= ((uint32_t)(key[4 * q + b] >> shift) & 255u) << (8 * b);
        }
      }
      __syncwarp();
      // Ranking: lanes of a round that hold the same digit must learn of each other (their ranks follow the lane order: the
      // sort is stable).  MATCH.ANY does that in hardware by looping over the distinct values of the warp -- ~30 of them with
      // 8-bit digits (issue slots 12 % busy, profiles/r1_summary.md).  Instead every lane ORs its bit into a per-warp, per-digit
      // mask word in shared memory and reads the word back: one shared atomic and two loads per round.
#pragma unroll
      for (int r = 0; r < RS_ITEMS; ++r) {
        const bool ok = (wl + 32 * r) < ntile;
        const uint32_t dig = (dg[r >> 2] >> (8 * (r & 3))) & 255u;
        // mask (low word) and running count (high word) of the warp's digit share one 64-bit word: one load, one store
        volatile unsigned long long* mc = &wmc[warp][dig];
        if (ok) atomicOr(reinterpret_cast<uint32_t*>(&wmc[warp][dig]), 1u << lane);
        __syncwarp();
        const unsigned long long v = ok ? *mc : 0ull;
        const unsigned peers = (unsigned)v;
        const uint32_t prev = (uint32_t)(v >> 32);
        __syncwarp();
        if (ok && (peers & lt) == 0) *mc = (unsigned long long)(prev + __popc(peers)) << 32;   // the lowest lane of each digit group updates and clears
        __syncwarp();
        const uint32_t rank = prev + __popc(peers & lt);
        if (r & 1) rk[r >> 1] |= rank << 16; else rk[r >> 1] = rank;
      }
    }
    __syncthreads();
    if (have) {
      // thread d: the warps' counts of digit d, (block scan) the tile position where digit d starts, and back into wmc's count halves the
      // tile position where warp w's keys of digit d start
      uint32_t wc[RS_WARPS];
#pragma unroll
      for (int w = 0; w < RS_WARPS; ++w) { wc[w] = (uint32_t)(wmc[w][d] >> 32); run += wc[w]; }
      if (CHAINED) {
        agg[(size_t)tile * 256 + d] = run;
        if (run) atomicAdd(gt + (size_t)(tile / RS_G) * 256 + d, run);
      }
      uint32_t inc = run;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += v; }
      if (lane == 31) wtot[warp] = inc;
      __syncthreads();
      uint32_t before = 0;
#pragma unroll
      for (int w = 0; w < RS_WARPS; ++w) if (w < warp) before += wtot[w];
      start = before + inc - run;
      {
        uint32_t at = start;
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) { reinterpret_cast<uint32_t*>(&wmc[w][d])[1] = at; at += wc[w]; }
      }
      if (CHAINED && threadIdx.x == 0) {                 // after the scan's barrier: all 256 words are written / added
        rs_st_release(desc + tile, 1u);
        closes = atomicAdd(gcount + tile / RS_G, 1u);    // (the release above ordered the block's RED.ADDs before this one)
      }
    }
    // ---- the staged tile: its prefix, then out in runs
    if (ptile >= 0) {
      const int pgrp = ptile / RS_G, gi = ptile % RS_G;
      const long long pbase = (long long)ptile * RS_TILE;
      const int pn = (int)((n - pbase) < (long long)RS_TILE ? (n - pbase) : (long long)RS_TILE);
      if (!CHAINED) {
        gbase[d] = offsets[(size_t)d * nblocks + ptile] - pstart;
      } else {
        uint32_t pin = 0, gxs = 0;                           // counts of the group's earlier tiles; of all earlier groups
        if (gi > 0 || pgrp > 0) {
          const unsigned want = ((1u << gi) - 1u) | (pgrp > 0 ? 1u << 16 : 0u);
          unsigned spins = 0;
          for (;;) {
            uint32_t f = 0;
            if (lane < gi) f = rs_ld_flag(desc + ptile - 1 - lane);
            else if (lane == 16 && pgrp > 0) f = rs_ld_flag(gxf + pgrp);
            __syncwarp();
            if ((__ballot_sync(0xffffffffu, f != 0u) & want) == want) break;
            if (++spins > RS_SPIN_LIMIT) { if (lane == 0) atomicExch(desc + nblocks + 1, 1u); __trap(); }
            __nanosleep(40);
          }
          uint32_t va[RS_G - 1];
#pragma unroll
          for (int q = 0; q < RS_G - 1; ++q) va[q] = q < gi ? __ldcg(agg + (size_t)(ptile - 1 - q) * 256 + d) : 0u;
          if (pgrp > 0) gxs = __ldcg(gx + (size_t)pgrp * 256 + d);
#pragma unroll
          for (int q = 0; q < RS_G - 1; ++q) pin += va[q];
        }
        gbase[d] = offsets[d] + gxs + pin - pstart;
      }
      __syncthreads();
#pragma unroll
      for (int r = 0; r < RS_ITEMS; ++r) {
        const int e = r * RS_THREADS + threadIdx.x;
        if (e < pn) {
          const K k = stage[e];
          const uint32_t pos = gbase[(uint32_t)(k >> shift) & 255u] + (uint32_t)e;
          keys_out[pos] = UNMAKE ? unkey_bits(k) : k;
          idx_out[pos] = ibuf[e];
          if (Pay2::ON) pay2_out[pos] = ibuf2[e];
        }
      }
    }
    if (!have) break;
    if (CHAINED && threadIdx.x == 0) wtot[RS_WARPS + 1] = closes;
    __syncthreads();                                      // the staging buffers are free; the warps' digit bases are complete
    if (CHAINED && wtot[RS_WARPS + 1] == RS_G - 1 && tile / RS_G + 1 < ngroups) {
      // this block's tile completed its group: GX[G1] of the next group = GX[G1-k] + GT[G1-1] + .. + GT[G1-k], k >= 1 the
      // nearest group whose GX is published (GX[0] = 0 is not stored) with every group in between complete
      const int G1 = tile / RS_G + 1;
      uint32_t sum = 0;
      unsigned spins = 0;
      for (;;) {
        const int k = lane + 1, gk = G1 - k;
        uint32_t fx = 0, fc = 0;
        if (k <= RS_GK && gk >= 0) {
          fx = gk == 0 ? 1u : rs_ld_flag(gxf + gk);
          fc = rs_ld_flag(gcount + gk) == (uint32_t)RS_G;
        }
        __syncwarp();
        const unsigned xm = __ballot_sync(0xffffffffu, fx != 0u), cm = __ballot_sync(0xffffffffu, fc != 0u);
        const int kx = xm ? __ffs(xm) : 0;                 // k of the nearest published prefix
        if (kx && (cm & ((1u << kx) - 1u)) == ((1u << kx) - 1u)) {
          uint32_t vg[RS_GK + 1];
          vg[0] = G1 - kx > 0 ? __ldcg(gx + (size_t)(G1 - kx) * 256 + d) : 0u;
#pragma unroll
          for (int q = 1; q <= RS_GK; ++q) vg[q] = q <= kx ? __ldcg(gt + (size_t)(G1 - q) * 256 + d) : 0u;
#pragma unroll
          for (int q = 0; q <= RS_GK; ++q) sum += vg[q];
          break;
        }
        if (++spins > RS_SPIN_LIMIT) { if (lane == 0) atomicExch(desc + nblocks + 1, 1u); __trap(); }
        __nanosleep(40);
      }
      gx[(size_t)G1 * 256 + d] = sum;
      __syncthreads();
      if (threadIdx.x == 0) rs_st_release(gxf + G1, 1u);
    }
    // the next tile's number: tickets are the look-back order, so it is asked for just before the tile starts (the staging
    // below hides the round trip)
    uint32_t next_ticket = (uint32_t)nblocks;
    if (CHAINED && threadIdx.x == 0) next_ticket = atomicAdd(desc + nblocks, 1u);
    // keys (again, from L2) and payload into digit order in shared memory (with a second payload: in two halves, for the registers)
    constexpr int EB = Pay2::ON ? RS_ITEMS / 2 : RS_ITEMS;
    auto cur = keys_in.at(tbase + wl);
#pragma unroll
    for (int r0 = 0; r0 < RS_ITEMS; r0 += EB) {
      K key[EB];
      uint32_t pay[EB];
      double p2[EB];
#pragma unroll
      for (int q = 0; q < EB; ++q) { key[q] = (wl + 32 * (r0 + q)) < ntile ? cur.again() : (K)0; cur.advance(32); }
#pragma unroll
      for (int q = 0; q < EB; ++q) {
        const int r = r0 + q;
        pay[q] = idx_in ? ((wl + 32 * r) < ntile ? __ldcg(idx_in + tbase + wl + 32 * r) : 0u) : (uint32_t)(tbase + wl + 32 * r);   // no payload array: the key's position
        p2[q] = (Pay2::ON && (wl + 32 * r) < ntile) ? pay2.at(tbase + wl + 32 * r, (uint32_t)key[q]) : 0.0;
      }
#pragma unroll
      for (int q = 0; q < EB; ++q) {
        const int r = r0 + q;
        if ((wl + 32 * r) < ntile) {
          const uint32_t dig = (dg[r >> 2] >> (8 * (r & 3))) & 255u;
          const uint32_t pos = reinterpret_cast<const uint32_t*>(&wmc[warp][dig])[1] + ((rk[r >> 1] >> (16 * (r & 1))) & 0xffffu);   // tile position
          stage[pos] = key[q];
          ibuf[pos] = pay[q];
          if (Pay2::ON) ibuf2[pos] = p2[q];
        }
      }
    }
    ptile = tile; prun = run; pstart = start;
    if (threadIdx.x == 0) wtot[RS_WARPS] = next_ticket;
    __syncthreads();
    tile = (int)wtot[RS_WARPS];
  }
}
template <class K, bool CHAINED, class In = PtrKeys<K>, bool UNMAKE = false, class Pay2 = NoPay2>
int rs_launch_scatter(const In kin, const uint32_t* iin, K* kout, uint32_t* iout, long long n, int shift, const uint32_t* offsets, int nblocks,
                      uint32_t* desc, cudaStream_t st, const Pay2 pay2 = Pay2(), double* pay2_out = nullptr) {
  static bool configured[RS_MAX_DEVICES] = {};
  int dev = 0;
  OCB_CUDA(cudaGetDevice(&dev));
  constexpr size_t smem = rs_scatter_smem<K, Pay2>();
  if (dev < 0 || dev >= RS_MAX_DEVICES || !configured[dev]) {
    OCB_CUDA(cudaFuncSetAttribute(rs_scatter<K, CHAINED, In, UNMAKE, Pay2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (dev >= 0 && dev < RS_MAX_DEVICES) configured[dev] = true;
  }
  int grid = nblocks;
  if (CHAINED) {                                         // persistent: exactly the resident blocks
    static int per_sm[RS_MAX_DEVICES] = {};
    const int di = dev >= 0 && dev < RS_MAX_DEVICES ? dev : 0;
    if (per_sm[di] <= 0) {
      int v = 0;
      OCB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, rs_scatter<K, CHAINED, In, UNMAKE, Pay2>, RS_THREADS, smem));
      per_sm[di] = v > 0 ? v : 1;
    }
    const int resident = per_sm[di] * ocb_sm_count();
    if (grid > resident) grid = resident;
  }
  rs_scatter<K, CHAINED, In, UNMAKE, Pay2><<<grid, RS_THREADS, smem, st>>>(kin, iin, kout, iout, n, shift, offsets, nblocks, desc, pay2, pay2_out);
  return OCB_OK;
}

// all digit histograms of the keys in one read: hist[pass][256] (global digit counts do not depend on the order of the keys)
template <class K, class In>
__global__ void __launch_bounds__(RS_THREADS, 4) rs_histogram_all(const In keys, long long n, uint32_t* __restrict__ hist) {
  constexpr int NP = KeyTraits<K>::BITS / 8;
  constexpr int HB = 8;                               // loads in flight per thread
  __shared__ uint32_t h[NP][256];
  for (int p = 0; p < NP; ++p) h[p][threadIdx.x] = 0;
  __syncthreads();
  for (long long base = (long long)blockIdx.x * (HB * RS_THREADS); base < n; base += (long long)gridDim.x * (HB * RS_THREADS)) {
    K keyr[HB];
    auto cur = keys.at(base + threadIdx.x < n ? base + threadIdx.x : 0);
#pragma unroll
    for (int r = 0; r < HB; ++r) {
      const long long t = base + r * RS_THREADS + threadIdx.x;
      keyr[r] = t < n ? cur.first() : (K)0;
      cur.advance(RS_THREADS);
    }
#pragma unroll
    for (int r = 0; r < HB; ++r) {
      const long long t = base + r * RS_THREADS + threadIdx.x;
      const bool ok = t < n;
      const K key = keyr[r];
      // the upper digits of real data are nearly constant: one add per warp when all its lanes agree on a digit (the bits in which
      // a lane differs from lane 0, one vote per digit)
      const unsigned act = __ballot_sync(0xffffffffu, ok);
      const K diff = key ^ __shfl_sync(0xffffffffu, key, 0);
#pragma unroll
      for (int p = 0; p < NP; ++p) {
        const uint32_t dig = (uint32_t)(key >> (8 * p)) & 255u;
        const bool same = __all_sync(0xffffffffu, ((uint32_t)(diff >> (8 * p)) & 255u) == 0u) && act == 0xffffffffu;
        if (same) { if ((threadIdx.x & 31) == 0) atomicAdd(&h[p][dig], 32u); }
        else if (ok) atomicAdd(&h[p][dig], 1u);
      }
    }
  }
  __syncthreads();
  for (int p = 0; p < NP; ++p) { const uint32_t c = h[p][threadIdx.x]; if (c) atomicAdd(hist + p * 256 + threadIdx.x, c); }
}
// exclusive scan of each pass's 256 counts, in place (one block per pass)
__global__ void __launch_bounds__(256) rs_digit_bases(uint32_t* __restrict__ hist) {
  __shared__ uint32_t sh[256];
  uint32_t* h = hist + blockIdx.x * 256;
  const uint32_t c = h[threadIdx.x];
  sh[threadIdx.x] = c;
  __syncthreads();
  for (int o = 1; o < 256; o <<= 1) {
    const uint32_t v = threadIdx.x >= o ? sh[threadIdx.x - o] : 0u;
    __syncthreads();
    sh[threadIdx.x] += v;
    __syncthreads();
  }
  h[threadIdx.x] = sh[threadIdx.x] - c;
}

// digit bases of the integers 0 .. n-1 (a permutation's values) for the digit at bits [shift, shift + 8), n <= 2^(shift + 8):
// digit d holds min(max(n - d 2^shift, 0), 2^shift) of them -- no histogram pass
__global__ void __launch_bounds__(256) rs_iota_digit_bases(long long n, int shift, uint32_t* __restrict__ bases) {
  const long long lo = (long long)threadIdx.x << shift;
  bases[threadIdx.x] = (uint32_t)(lo < n ? lo : n);
}

// ---- block scans ----------------------------------------------------------------------------------------------
enum { OP_ADD = 0, OP_MAX = 1, OP_MIN = 2 };
template <class T, int OP> __device__ __forceinline__ T sc_op(T a, T b) { return OP == OP_ADD ? a + b : (OP == OP_MAX ? (a > b ? a : b) : (a < b ? a : b)); }
template <class T, int OP> __device__ __forceinline__ T sc_identity();
template <> __device__ __forceinline__ uint32_t sc_identity<uint32_t, OP_ADD>() { return 0u; }
template <> __device__ __forceinline__ double sc_identity<double, OP_ADD>() { return 0.0; }
template <> __device__ __forceinline__ double sc_identity<double, OP_MAX>() { return -DBL_MAX; }
template <> __device__ __forceinline__ double sc_identity<double, OP_MIN>() { return DBL_MAX; }

constexpr int SC_THREADS = 256, SC_ITEMS = 8, SC_CHUNK = SC_THREADS * SC_ITEMS;

// inclusive scan of one chunk held as SC_ITEMS consecutive items per thread; returns the chunk total
template <class T, int OP>
__device__ __forceinline__ T block_scan_chunk(T (&x)[SC_ITEMS], T* sh) {
  T run = x[0];
#pragma unroll
  for (int q = 1; q < SC_ITEMS; ++q) { run = sc_op<T, OP>(run, x[q]); x[q] = run; }
  // scan of the per-thread totals: warp shuffle, then across warps
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  T v = run;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { T u = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v = sc_op<T, OP>(u, v); }
  if (lane == 31) sh[warp] = v;
  __syncthreads();
  if (warp == 0) {
    T wv = lane < SC_THREADS / 32 ? sh[lane] : sc_identity<T, OP>();
#pragma unroll
    for (int o = 1; o < SC_THREADS / 32; o <<= 1) { T u = __shfl_up_sync(0xffffffffu, wv, o); if (lane >= o) wv = sc_op<T, OP>(u, wv); }
    if (lane < SC_THREADS / 32) sh[lane] = wv;
  }
  __syncthreads();
  const T total = sh[SC_THREADS / 32 - 1];
  T excl = sc_identity<T, OP>();                 // exclusive prefix of this thread
  if (warp > 0) excl = sh[warp - 1];
  T up = __shfl_up_sync(0xffffffffu, v, 1);
  if (lane > 0) excl = sc_op<T, OP>(excl, up);
#pragma unroll
  for (int q = 0; q < SC_ITEMS; ++q) x[q] = sc_op<T, OP>(excl, x[q]);
  __syncthreads();
  return total;
}

// scan inputs: an array, or values formed on the fly from other arrays (in(t))
template <class T> struct PtrInput {
  const T* p;
  __device__ __forceinline__ T operator()(long long t) const { return p[t]; }
};

// A thread scans SC_ITEMS consecutive elements, but global memory is read and written element q * SC_THREADS + thread at a time
// (consecutive lanes, consecutive addresses) and the chunk is transposed through shared memory -- one spare element per thread's
// run keeps both access patterns off each other's banks.  (Reading the runs directly costs a 64-byte lane stride per load and
// quarter-written sectors per store: the Psi scan ran at 2.6 TB/s.)
constexpr int SC_TILE = SC_CHUNK + SC_THREADS;
template <class T, int OP, bool REV, class In>
__device__ __forceinline__ void sc_load_runs(const In& in, long long n, long long chunk_base, T* tile, T (&x)[SC_ITEMS]) {
#pragma unroll
  for (int q = 0; q < SC_ITEMS; ++q) {
    const int e = q * SC_THREADS + threadIdx.x;
    const long long t = chunk_base + e;
    tile[e + e / SC_ITEMS] = t < n ? in(REV ? n - 1 - t : t) : sc_identity<T, OP>();
  }
  __syncthreads();
#pragma unroll
  for (int q = 0; q < SC_ITEMS; ++q) x[q] = tile[threadIdx.x * (SC_ITEMS + 1) + q];
}
// phase 1: chunk totals.  REV: the scan runs from the end of the array (element n-1 first).  (Same arrangement and scan as
// phase 3, so that a chunk's total is bit for bit the last inclusive value phase 3 forms in it.)
template <class T, int OP, bool REV, class In>
__global__ void __launch_bounds__(SC_THREADS) sc_chunk_totals(const In in, long long n, T* __restrict__ totals) {
  __shared__ T sh[SC_THREADS / 32];
  __shared__ T tile[SC_TILE];
  T x[SC_ITEMS];
  sc_load_runs<T, OP, REV, In>(in, n, (long long)blockIdx.x * SC_CHUNK, tile, x);
  const T total = block_scan_chunk<T, OP>(x, sh);
  if (threadIdx.x == 0) totals[blockIdx.x] = total;
}
// phase 2: exclusive scan of the chunk totals by one block (loops when there are more than SC_CHUNK chunks)
template <class T, int OP>
__global__ void __launch_bounds__(SC_THREADS) sc_scan_totals(T* __restrict__ totals, int nchunks) {
  __shared__ T sh[SC_THREADS / 32];
  __shared__ T last[SC_THREADS];
  T carry = sc_identity<T, OP>();
  for (long long base0 = 0; base0 < nchunks; base0 += SC_CHUNK) {
    T x[SC_ITEMS];
    const long long base = base0 + (long long)threadIdx.x * SC_ITEMS;
#pragma unroll
    for (int q = 0; q < SC_ITEMS; ++q) { const long long t = base + q; x[q] = t < nchunks ? totals[t] : sc_identity<T, OP>(); }
    const T total = block_scan_chunk<T, OP>(x, sh);
    last[threadIdx.x] = x[SC_ITEMS - 1];
    __syncthreads();
#pragma unroll
    for (int q = 0; q < SC_ITEMS; ++q) {
      const long long t = base + q;
      if (t < nchunks) {   // exclusive value = carry (op) inclusive value of the previous element
        const T prev = q > 0 ? x[q - 1] : (threadIdx.x > 0 ? last[threadIdx.x - 1] : sc_identity<T, OP>());
        totals[t] = sc_op<T, OP>(carry, prev);
      }
    }
    carry = sc_op<T, OP>(carry, total);
    __syncthreads();
  }
}
// phase 3: scan each chunk and combine with its prefix.  EXCL: write the exclusive scan.
template <class T, int OP, bool REV, bool EXCL, class In>
__global__ void __launch_bounds__(SC_THREADS) sc_apply(const In in, long long n, const T* __restrict__ totals, T* __restrict__ out) {
  __shared__ T sh[SC_THREADS / 32];
  __shared__ T last[SC_THREADS];
  __shared__ T tile[SC_TILE];
  T x[SC_ITEMS];
  const long long chunk_base = (long long)blockIdx.x * SC_CHUNK;
  sc_load_runs<T, OP, REV, In>(in, n, chunk_base, tile, x);
  block_scan_chunk<T, OP>(x, sh);                       // (ends with a barrier: the tile can be overwritten)
  const T prefix = totals[blockIdx.x];
  if (EXCL) { last[threadIdx.x] = x[SC_ITEMS - 1]; __syncthreads(); }
#pragma unroll
  for (int q = 0; q < SC_ITEMS; ++q) {
    T v;
    if (EXCL) {
      T prev = q > 0 ? x[q - 1] : (threadIdx.x > 0 ? last[threadIdx.x - 1] : sc_identity<T, OP>());
      v = sc_op<T, OP>(prefix, prev);
    } else {
      v = sc_op<T, OP>(prefix, x[q]);
    }
    tile[threadIdx.x * (SC_ITEMS + 1) + q] = v;
  }
  __syncthreads();
#pragma unroll
  for (int q = 0; q < SC_ITEMS; ++q) {
    const int e = q * SC_THREADS + threadIdx.x;
    const long long t = chunk_base + e;
    if (t < n) out[REV ? n - 1 - t : t] = tile[e + e / SC_ITEMS];
  }
}

template <class T, int OP, bool REV, bool EXCL, class In>
int device_scan_of(const In in, T* out, long long n, T* totals_scratch, cudaStream_t st) {
  const int nchunks = (int)((n + SC_CHUNK - 1) / SC_CHUNK);
  sc_chunk_totals<T, OP, REV, In><<<nchunks, SC_THREADS, 0, st>>>(in, n, totals_scratch);
  sc_scan_totals<T, OP><<<1, SC_THREADS, 0, st>>>(totals_scratch, nchunks);
  sc_apply<T, OP, REV, EXCL, In><<<nchunks, SC_THREADS, 0, st>>>(in, n, totals_scratch, out);
  OCB_CUDA(cudaGetLastError());
  return OCB_OK;
}
template <class T, int OP, bool REV, bool EXCL>
int device_scan(const T* in, T* out, long long n, T* totals_scratch, cudaStream_t st) {
  return device_scan_of<T, OP, REV, EXCL, PtrInput<T>>(PtrInput<T>{in}, out, n, totals_scratch, st);
}

// Sort of n (key, position) pairs.  `first` is where the keys come from (the first pass makes them on the fly and its payload is
// the position).  Passes alternate between the (a) and (b) scratch pairs; the LAST pass writes to (final_keys, final_idx) when
// they are given -- the caller's own arrays -- and stores values instead of keys with `unmake`.  *keys_sorted / *idx_sorted are
// where the result is.
template <class K, class In>
int radix_sort_pairs(const In first, K* keys_a, uint32_t* idx_a, K* keys_b, uint32_t* idx_b, long long n, uint32_t* table, uint32_t* table_tot,
                     int nblocks, cudaStream_t st, K** keys_sorted, uint32_t** idx_sorted, bool unmake = false, K* final_keys = nullptr,
                     uint32_t* final_idx = nullptr) {
  constexpr int NP = KeyTraits<K>::BITS / 8;
  const long long ntab = 256ll * nblocks;
  if (n < (1ll << 30) && !getenv("OCB_SORT_TABLE_SCAN")) {
    // onesweep: one histogram of all digits up front, then ONE kernel per digit.  `table` (rs_table_words) holds the tile / group
    // descriptors of the current pass; `table_tot` (sized for it by the callers) the digit histograms / bases.
    uint32_t* hist = table_tot;
    OCB_CUDA(cudaMemsetAsync(hist, 0, sizeof(uint32_t) * NP * 256, st));
    int per_sm = 0;                                     // exactly the resident blocks: a partial second wave would cost a whole one
    OCB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, rs_histogram_all<K, In>, RS_THREADS, 0));
    int hb = ocb_sm_count() * (per_sm > 0 ? per_sm : 1);
    if (hb > 2 * nblocks) hb = 2 * nblocks;
    rs_histogram_all<K, In><<<hb, RS_THREADS, 0, st>>>(first, n, hist);
    rs_digit_bases<<<NP, 256, 0, st>>>(hist);
    K* kin = nullptr; uint32_t* iin = nullptr;
    for (int p = 0; p < NP; ++p) {
      const bool last = p == NP - 1;
      K* kout = last && final_keys ? final_keys : ((p & 1) ? keys_b : keys_a);
      uint32_t* iout = last && final_idx ? final_idx : ((p & 1) ? idx_b : idx_a);
      OCB_CUDA(cudaMemsetAsync(table, 0, sizeof(uint32_t) * rs_desc_clear(nblocks), st));      // flags, ticket, error word, group totals
      int rc;
      if (p == 0) rc = rs_launch_scatter<K, true, In, false>(first, nullptr, kout, iout, n, 0, hist, nblocks, table, st);
      else if (last && unmake) rc = rs_launch_scatter<K, true, PtrKeys<K>, true>(PtrKeys<K>{kin}, iin, kout, iout, n, 8 * p, hist + p * 256, nblocks, table, st);
      else rc = rs_launch_scatter<K, true, PtrKeys<K>, false>(PtrKeys<K>{kin}, iin, kout, iout, n, 8 * p, hist + p * 256, nblocks, table, st);
      if (rc != OCB_OK) return rc;
      kin = kout; iin = iout;
    }
    OCB_CUDA(cudaGetLastError());
    *keys_sorted = kin; *idx_sorted = iin;
    return OCB_OK;
  }
  // per-pass histogram table + scan (n >= 2^30, or OCB_SORT_TABLE_SCAN): keys made by a pass of their own, values by another
  K* kin = keys_a; K* kout = keys_b;
  uint32_t* iin = idx_a; uint32_t* iout = idx_b;
  rs_materialize_keys<K, In><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(first, kin, iin, n);
  for (int shift = 0; shift < KeyTraits<K>::BITS; shift += 8) {
    rs_histogram<K><<<nblocks, RS_THREADS, 0, st>>>(kin, n, shift, table, nblocks);
    int rc = device_scan<uint32_t, OP_ADD, false, true>(table, table, ntab, table_tot, st);
    if (rc != OCB_OK) return rc;
    rc = rs_launch_scatter<K, false>(PtrKeys<K>{kin}, iin, kout, iout, n, shift, table, nblocks, nullptr, st);
    if (rc != OCB_OK) return rc;
    K* tk = kin; kin = kout; kout = tk;
    uint32_t* ti = iin; iin = iout; iout = ti;
  }
  if (unmake) {
    typename KeyTraits<K>::F* dst = reinterpret_cast<typename KeyTraits<K>::F*>(final_keys ? final_keys : kout);
    rs_unmake_keys<K><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(kin, dst, n);
    kin = reinterpret_cast<K*>(dst);
  } else if (final_keys) {
    OCB_CUDA(cudaMemcpyAsync(final_keys, kin, n * sizeof(K), cudaMemcpyDeviceToDevice, st));
    kin = final_keys;
  }
  if (final_idx) {
    OCB_CUDA(cudaMemcpyAsync(final_idx, iin, n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
    iin = final_idx;
  }
  OCB_CUDA(cudaGetLastError());
  *keys_sorted = kin; *idx_sorted = iin;
  return OCB_OK;
}

// ---- BPE epilogue kernels ---------------------------------------------------------------------------------------
template <class T>
__global__ void bpe_gather_interior(const T* __restrict__ b, long long s0, long long s1, long long s2, int Nx, int Ny, long long n, T* __restrict__ work) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;   // b pre-offset to index (1,1,1) -> element 0
  if (t >= n) return;
  const int i = (int)(t % Nx), j = (int)((t / Nx) % Ny), k = (int)(t / ((long long)Nx * Ny));
  work[t] = b[i * s0 + j * s1 + k * s2];
}
// sorted cell volumes: dV[n] = dx*dy*dzc[k(perm[n])]
template <class T>
__global__ void bpe_sorted_volumes(const uint32_t* __restrict__ perm, long long n, int Nx, int Ny, T dxdy, const T* __restrict__ dzc1, T dz_uniform,
                                   double* __restrict__ dV) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= n) return;
  const long long k = perm[t] / ((long long)Nx * Ny);
  const T dz = dzc1 ? dzc1[k] : dz_uniform;
  dV[t] = (double)(dxdy * dz);
}
// ThreeDimensionalSort slot centres: z* = z_bot + (cum - dV/2)/A ; also b*dz for the Ψ scan
template <class T>
__global__ void bpe_slot_centres(const double* __restrict__ cum, const double* __restrict__ dV, const T* __restrict__ bs, long long n, double z_bot, double A,
                                 T* __restrict__ zs_sorted, double* __restrict__ bdz) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= n) return;
  if (zs_sorted) zs_sorted[t] = (T)(z_bot + (cum[t] - dV[t] / 2) / A);
  // faces[t] = z_bot + cum[t-1]/A, faces[t+1] = z_bot + cum[t]/A  (slot_faces): dz = their difference
  const double f0 = z_bot + (t > 0 ? cum[t - 1] : 0.0) / A, f1 = z_bot + cum[t] / A;
  bdz[t] = (double)bs[t] * (f1 - f0);
}
// HeavisideIntegral: run starts / ends of equal keys -> masked `below` and `cum` for the max / reversed-min scans
template <class T>
__global__ void bpe_heaviside_prepare(const T* __restrict__ bs, const double* __restrict__ cum, const double* __restrict__ dV, long long n,
                                      double* __restrict__ den_in, double* __restrict__ nol_in) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= n) return;
  const bool start = t == 0 || bs[t] != bs[t - 1];
  const bool end = t == n - 1 || bs[t + 1] != bs[t];
  den_in[t] = start ? cum[t] - dV[t] : -DBL_MAX;
  nol_in[t] = end ? cum[t] : DBL_MAX;
}
template <class T>
__global__ void bpe_heaviside_centres(const double* __restrict__ denser, const double* __restrict__ no_lighter, long long n, double z_bot, double A,
                                      T* __restrict__ zs_sorted) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t < n) zs_sorted[t] = (T)(z_bot + (denser[t] + no_lighter[t]) / (2 * A));
}
// Ψ(ζ) = Ψface[k] + b*[k] (ζ - faces[k]),  k = clamp(searchsortedlast(faces, ζ), 1, N)  (1-based), faces[m] = z_bot + cum[m-1]/A
template <class T>
__device__ __forceinline__ double bpe_psi(double zeta, const double* cum, const double* psi_incl, const T* bs, long long n, double z_bot, double A) {
  // faces has n+1 entries f[0..n] (0-based), f[m] = z_bot + (m ? cum[m-1] : 0)/A ; last index with f[m] <= zeta
  long long lo = -1, hi = n + 1;      // invariant: f[lo] <= zeta < f[hi]  (f[-1] = -inf, f[n+1] = +inf)
  while (hi - lo > 1) {
    const long long mid = (lo + hi) >> 1;
    const double f = z_bot + (mid ? cum[mid - 1] : 0.0) / A;
    if (f <= zeta) lo = mid; else hi = mid;
  }
  long long k = lo + 1;               // 1-based searchsortedlast
  k = k < 1 ? 1 : (k > n ? n : k);
  const double fk = z_bot + (k - 1 ? cum[k - 2] : 0.0) / A;
  const double psi_face = k - 1 ? psi_incl[k - 2] : 0.0;
  return psi_face + (double)bs[k - 1] * (zeta - fk);
}
// Ψ at the Nz cell-centre heights: z of a cell depends on k only, so the N per-cell binary searches of the
// reference (reference_potential_function evaluated at every z, :359-368 / :383-397) collapse to Nz searches.
template <class T>
__global__ void bpe_psi_levels(const double* __restrict__ cum, const double* __restrict__ psi_incl, const T* __restrict__ bs, long long n, int Nz,
                               double z_bot, double A, const T* __restrict__ zc1, T z0, T dz_uniform, double* __restrict__ psi_level) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= Nz) return;
  const double zcell = zc1 ? (double)zc1[k] : (double)z0 + (double)dz_uniform * (k + 0.5);
  psi_level[k] = bpe_psi<T>(zcell, cum, psi_incl, bs, n, z_bot, A);
}
// scatter z* and Ψδ back through the permutation into the destination fields.  SLOT_KNOWN (ThreeDimensionalSort): the
// t-th sorted cell sits in slot t, so Ψ(z*) needs no search: Ψface[t] + b*[t] (z* - faces[t]).
template <class T, bool SLOT_KNOWN>
__global__ void bpe_scatter(const uint32_t* __restrict__ perm, const T* __restrict__ zs_sorted, const double* __restrict__ cum, const double* __restrict__ psi_incl,
                            const T* __restrict__ bs, long long n, int Nx, int Ny, double z_bot, double A, const double* __restrict__ psi_level,
                            T* __restrict__ zstar, long long zs0, long long zs1, long long zs2, T* __restrict__ psid, long long ps0, long long ps1, long long ps2) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= n) return;
  const uint32_t cell = perm[t];
  const int i = (int)(cell % Nx), j = (int)((cell / Nx) % Ny), k = (int)(cell / ((long long)Nx * Ny));
  const T zs = zs_sorted[t];
  zstar[i * zs0 + j * zs1 + k * zs2] = zs;
  if (psid) {
    double ps;
    if (SLOT_KNOWN) {
      const double ft = z_bot + (t ? cum[t - 1] : 0.0) / A;
      ps = (t ? psi_incl[t - 1] : 0.0) + (double)bs[t] * ((double)zs - ft);
    } else {
      ps = bpe_psi<T>((double)zs, cum, psi_incl, bs, n, z_bot, A);
    }
    psid[i * ps0 + j * ps1 + k * ps2] = (T)(psi_level[k] - ps);
  }
}
__global__ void bpe_perm_to_int64(const uint32_t* __restrict__ perm, long long n, int64_t* __restrict__ out) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t < n) out[t] = (int64_t)perm[t] + 1;
}

// ---- ThreeDimensionalSort on its (always uniform-volume) grid: the epilogue without the N-length temporaries -----------
// With equal cell volumes the cumulative volume of the first m sorted cells is m dV, so the volume array, its scan, the
// sorted-height array and the b* dz array need not exist: slot faces, slot centres and the Psi increments are formed where they
// are used, from the slot number and the sorted keys (same expressions as slot_centers! / slot_faces / fill_reference_potential!,
// BackgroundPotentialEnergyEquation.jl:326-348, 383-397).  Bytes per cell: 20 (keys from the field) + 24 (Psi scan) +
// 36 + the two scattered stores, instead of 232.
struct UniformSlots {          // faces[m] = z_bot + cum[m-1]/A with cum[m-1] = m dV  (0-based face m below slot m)
  double z_bot, A, dV;
  __host__ __device__ __forceinline__ double face(long long m) const { return z_bot + ((double)m * dV) / A; }
  __host__ __device__ __forceinline__ double centre(long long t) const { return z_bot + ((double)(t + 1) * dV - dV / 2) / A; }
};
template <class T, class K> struct BdzOfKeys {       // b*[t] (faces[t+1] - faces[t])
  const K* ks; UniformSlots U;
  __device__ __forceinline__ double operator()(long long t) const { return (double)value_of_bits(ks[t]) * (U.face(t + 1) - U.face(t)); }
};
template <class T, class K>
__global__ void bpe_psi_levels_uniform(const double* __restrict__ psi_incl, const K* __restrict__ ks, long long n, int Nz, UniformSlots U,
                                       T z0, T dz_uniform, double* __restrict__ psi_level) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= Nz) return;
  const double zeta = (double)z0 + (double)dz_uniform * (k + 0.5);
  long long lo = -1, hi = n + 1;      // searchsortedlast(faces, zeta), clamped to 1..n (reference_potential_function :359-368)
  while (hi - lo > 1) { const long long mid = (lo + hi) >> 1; if (U.face(mid) <= zeta) lo = mid; else hi = mid; }
  long long m = lo + 1;
  m = m < 1 ? 1 : (m > n ? n : m);
  psi_level[k] = (m - 1 ? psi_incl[m - 2] : 0.0) + (double)value_of_bits(ks[m - 1]) * (zeta - U.face(m - 1));
}
template <class T, class K>
__global__ void bpe_scatter_uniform(const uint32_t* __restrict__ perm, const K* __restrict__ ks, const double* __restrict__ psi_incl, long long n,
                                    int Nx, int Ny, UniformSlots U, const double* __restrict__ psi_level, T* __restrict__ zstar, long long zs0,
                                    long long zs1, long long zs2, T* __restrict__ psid, long long ps0, long long ps1, long long ps2,
                                    int64_t* __restrict__ perm_out, T* __restrict__ bs_out) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= n) return;
  const uint32_t cell = perm[t];
  const int i = (int)(cell % Nx), j = (int)((cell / Nx) % Ny), k = (int)(cell / ((long long)Nx * Ny));
  const T bs = value_of_bits(ks[t]);
  const T zs = (T)U.centre(t);
  zstar[i * zs0 + j * zs1 + k * zs2] = zs;
  if (psid) {
    const double ps = (t ? psi_incl[t - 1] : 0.0) + (double)bs * ((double)zs - U.face(t));
    psid[i * ps0 + j * ps1 + k * ps2] = (T)(psi_level[k] - ps);
  }
  if (perm_out) perm_out[t] = (int64_t)cell + 1;
  if (bs_out) bs_out[t] = bs;
}

// The same scatter, fed in DESTINATION order: (cell, rank) pairs grouped by the region of the cell (one stable radix pass on
// the upper bits of the cell index).  The two 8-byte stores of a block then land in a few megabytes of the destination fields,
// which stay in L2 until their sectors are complete, instead of 2 N read-modify-written sectors all over the arrays; the
// price is two gathers (sorted key and Psi prefix at the rank), which are reads.
template <class T, class K>
__global__ void bpe_scatter_bucketed(const uint32_t* __restrict__ cells, const uint32_t* __restrict__ ranks, const K* __restrict__ ks,
                                     const double* __restrict__ psi_incl, long long n, int Nx, int Ny, UniformSlots U,
                                     const double* __restrict__ psi_level, T* __restrict__ zstar, long long zs0, long long zs1, long long zs2,
                                     T* __restrict__ psid, long long ps0, long long ps1, long long ps2) {
  const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e >= n) return;
  const uint32_t cell = cells[e];
  const long long t = ranks[e];
  const int i = (int)(cell % Nx), j = (int)((cell / Nx) % Ny), k = (int)(cell / ((long long)Nx * Ny));
  const T zs = (T)U.centre(t);
  zstar[i * zs0 + j * zs1 + k * zs2] = zs;
  if (psid) {
    const T bs = value_of_bits(__ldg(ks + t));
    const double ps = (t ? __ldg(psi_incl + t - 1) : 0.0) + (double)bs * ((double)zs - U.face(t));
    psid[i * ps0 + j * ps1 + k * ps2] = (T)(psi_level[k] - ps);
  }
}
// The reference potential of the cell of rank t, Psi(z*) (fill_reference_potential!), formed where the sorted keys and the
// running sum are read in order: the bucketing pass carries it with the (cell, rank) pair.
template <class T, class K> struct PsiOfRank {
  static constexpr bool ON = true;
  const K* ks; const double* psi_incl; UniformSlots U; int64_t* perm_out;      // ks: the sorted values (bit patterns)
  __device__ __forceinline__ double at(long long t, uint32_t cell) const {
    if (perm_out) perm_out[t] = (int64_t)cell + 1;                             // the 1-based sortperm, written where it is read in order
    const T bs = value_of_bits(__ldg(ks + t));
    const T zs = (T)U.centre(t);
    return (t ? __ldg(psi_incl + t - 1) : 0.0) + (double)bs * ((double)zs - U.face(t));
  }
};
// final scatter of (z*, Psi delta) from the bucketed (cell, rank, Psi) triples: every read in order, the stores within one region
template <class T>
__global__ void bpe_scatter_triples(const uint32_t* __restrict__ cells, const uint32_t* __restrict__ ranks, const double* __restrict__ psis,
                                    long long n, FastDiv dx, FastDiv dy, UniformSlots U, const double* __restrict__ psi_level,
                                    T* __restrict__ zstar, long long zs0, long long zs1, long long zs2,
                                    T* __restrict__ psid, long long ps0, long long ps1, long long ps2) {
  const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e >= n) return;
  const uint32_t cell = __ldg(cells + e);
  const long long t = __ldg(ranks + e);
  const uint32_t q = dx.div(cell), i = cell - q * dx.d, k = dy.div(q), j = q - k * dy.d;
  zstar[i * zs0 + j * zs1 + k * zs2] = (T)U.centre(t);
  psid[i * ps0 + j * ps1 + k * ps2] = (T)(psi_level[k] - __ldg(psis + e));
}
template <class T, class K>
__global__ void bpe_sorted_outputs(const uint32_t* __restrict__ perm, const K* __restrict__ ks, long long n, int64_t* __restrict__ perm_out,
                                   T* __restrict__ bs_out) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= n) return;
  if (perm_out) perm_out[t] = (int64_t)perm[t] + 1;
  if (bs_out) bs_out[t] = value_of_bits(ks[t]);
}

struct Scratch {
  cudaStream_t st;
  void* ptrs[24];
  int n = 0;
  explicit Scratch(cudaStream_t s) : st(s) {}
  template <class T> T* get(size_t count) {
    void* p = nullptr;
    if (cudaMallocAsync(&p, count * sizeof(T) + 256, st) != cudaSuccess) return nullptr;
    ptrs[n++] = p;
    return (T*)p;
  }
  ~Scratch() { for (int i = 0; i < n; ++i) cudaFreeAsync(ptrs[i], st); }
};

template <class T, class K>
int sort_impl(const T* keys_in, T* keys_out, uint32_t* perm_out, long long n, cudaStream_t st) {
  Scratch S(st);
  const int nblocks = (int)((n + RS_TILE - 1) / RS_TILE);
  K* ka = S.get<K>(n); K* kb = S.get<K>(n);
  uint32_t* ia = S.get<uint32_t>(n); uint32_t* ib = S.get<uint32_t>(n);
  uint32_t* table = S.get<uint32_t>(rs_table_words(nblocks));
  uint32_t* tot = S.get<uint32_t>((256ull * nblocks + SC_CHUNK - 1) / SC_CHUNK + 1 + 8 * 256);   // + the digit histograms of the onesweep path
  if (!ka || !kb || !ia || !ib || !table || !tot) OCB_FAIL(OCB_ERR_ALLOC, "out of device memory for the sort scratch");
  // the first pass reads the caller's values, the last one writes the sorted values and the permutation into the caller's arrays
  K* ks; uint32_t* is;
  int rc = radix_sort_pairs<K, FloatKeys<K>>(FloatKeys<K>{keys_in}, ka, ia, kb, ib, n, table, tot, nblocks, st, &ks, &is, true,
                                             reinterpret_cast<K*>(keys_out), perm_out);
  if (rc != OCB_OK) return rc;
  OCB_CUDA(cudaGetLastError());
  return OCB_OK;
}

template <class T, class K>
int bpe_sort3d_uniform(const ocb_grid* g, const ocb_field* b, const ocb_field* zstar, const ocb_field* psid, int64_t* perm_out, void* bs_out,
                       cudaStream_t st) {
  const int Nx = g->N[0], Ny = g->N[1], Nz = g->N[2];
  const long long n = (long long)Nx * Ny * Nz;
  Scratch S(st);
  const int nblocks = (int)((n + RS_TILE - 1) / RS_TILE);
  K* ka = S.get<K>(n); K* kb = S.get<K>(n);
  uint32_t* ia = S.get<uint32_t>(n); uint32_t* ib = S.get<uint32_t>(n);
  uint32_t* table = S.get<uint32_t>(rs_table_words(nblocks));
  uint32_t* tot = S.get<uint32_t>((256ull * nblocks + SC_CHUNK - 1) / SC_CHUNK + 1 + 8 * 256);   // + the digit histograms of the onesweep path
  double* psi = S.get<double>(n);
  double* dtot = S.get<double>((n + SC_CHUNK - 1) / SC_CHUNK + 1);
  double* psi_level = S.get<double>(Nz);
  if (!ka || !kb || !ia || !ib || !table || !tot || !psi || !dtot || !psi_level)
    OCB_FAIL(OCB_ERR_ALLOC, "out of device memory for the reference-height scratch");
  HFld<T> B, Z, P;
  int rc = make_fld<T>(*b, &B, "b", false); if (rc != OCB_OK) return rc;
  rc = make_fld<T>(*zstar, &Z, "zstar", false); if (rc != OCB_OK) return rc;
  OCB_REQUIRE(B.p && Z.p, "b and zstar must be array fields");
  P.p = nullptr; P.sx = P.sy = P.sz = 0;
  if (psid && psid->data) { rc = make_fld<T>(*psid, &P, "psi_delta", false); if (rc != OCB_OK) return rc; }
  const unsigned gr = (unsigned)((n + 255) / 256);
  // keys straight from the field (the first pass of the sort makes them); the last pass stores the sorted VALUES -- into the
  // caller's sorted-buoyancy array when there is one -- and the epilogue reads those
  K* ks; uint32_t* perm;
  rc = radix_sort_pairs<K, FieldKeys<T, K>>(FieldKeys<T, K>{B.p + (B.sx + B.sy + B.sz), B.sx, B.sy, B.sz, FastDiv((uint32_t)Nx), FastDiv((uint32_t)Ny)}, ka, ia, kb, ib, n, table, tot,
                                            nblocks, st, &ks, &perm, true, reinterpret_cast<K*>(bs_out));
  const bool bs_done = bs_out && (void*)ks == bs_out;
  if (rc != OCB_OK) return rc;
  UniformSlots U;
  U.z_bot = g->z_bottom; U.A = (g->L[0] * g->L[1] * g->L[2]) / g->L[2];
  U.dV = (double)((T)(g->d[0] * g->d[1]) * (T)g->d[2]);               // the cell volume in the grid's precision, as the general path forms it
  T* p1 = P.p ? (T*)P.p + (P.sx + P.sy + P.sz) : nullptr;
  if (p1) {
    rc = device_scan_of<double, OP_ADD, false, false, BdzOfKeys<T, K>>(BdzOfKeys<T, K>{ks, U}, psi, n, dtot, st);
    if (rc != OCB_OK) return rc;
    bpe_psi_levels_uniform<T, K><<<(Nz + 127) / 128, 128, 0, st>>>(psi, ks, n, Nz, U, (T)g->z_bottom, (T)g->d[2], psi_level);
  }
  T* z1 = (T*)Z.p + (Z.sx + Z.sy + Z.sz);
  if ((n >= (1ll << 20) || getenv("OCB_BPE_BUCKETED")) && !getenv("OCB_BPE_DIRECT_SCATTER")) {
    // one stable radix pass on bits [shift, shift + 8) of the cell index: 256 destination regions (2 MB of each field at 67 M cells).
    // The sort's idle ping-pong buffer holds the bucketed pairs.
    int bits = 0;
    while ((1ll << bits) < n) ++bits;
    const int shift = bits > 8 ? bits - 8 : 0;
    uint32_t* cells2 = reinterpret_cast<uint32_t*>(ks == ka ? kb : ka);        // (the last pass read ka and wrote ks: both scratch pairs are idle)
    uint32_t* ranks2 = (perm == ia) ? ib : ia;
    double* psis2 = (p1 && !getenv("OCB_BPE_GATHER_PSI")) ? S.get<double>(n) : nullptr;     // no room: the gathering scatter
    // the values of a permutation: the digit bases are known without a histogram, so the pass is one chained scatter
    uint32_t* bases = tot;
    rs_iota_digit_bases<<<1, 256, 0, st>>>(n, shift, bases);
    OCB_CUDA(cudaMemsetAsync(table, 0, sizeof(uint32_t) * rs_desc_clear(nblocks), st));
    if (p1 && psis2) {
      // with Psi: the pass carries Psi(rank) as a second payload, so the final scatter reads everything in order (the gathers of the
      // sorted key and the running sum at the rank, a 32-byte sector each, were 2.3 ms of the 8.7 at 67 M cells)
      const int rc2 = rs_launch_scatter<uint32_t, true, PtrKeys<uint32_t>, false, PsiOfRank<T, K>>(
          PtrKeys<uint32_t>{perm}, nullptr, cells2, ranks2, n, shift, bases, nblocks, table, st, PsiOfRank<T, K>{ks, psi, U, perm_out}, psis2);
      if (rc2 != OCB_OK) return rc2;
      bpe_scatter_triples<T><<<gr, 256, 0, st>>>(cells2, ranks2, psis2, n, FastDiv((uint32_t)Nx), FastDiv((uint32_t)Ny), U, psi_level, z1, Z.sx, Z.sy,
                                                 Z.sz, p1, P.sx, P.sy, P.sz);
    } else {
      const int rc2 = rs_launch_scatter<uint32_t, true>(PtrKeys<uint32_t>{perm}, nullptr, cells2, ranks2, n, shift, bases, nblocks, table, st);
      if (rc2 != OCB_OK) return rc2;
      bpe_scatter_bucketed<T, K><<<gr, 256, 0, st>>>(cells2, ranks2, ks, psi, n, Nx, Ny, U, psi_level, z1, Z.sx, Z.sy, Z.sz, p1, P.sx, P.sy, P.sz);
    }
    {
      int64_t* perm_left = (p1 && psis2) ? nullptr : perm_out;                  // (written by the bucketing pass above)
      T* bs_left = bs_done ? nullptr : (T*)bs_out;
      if (perm_left || bs_left) bpe_sorted_outputs<T, K><<<gr, 256, 0, st>>>(perm, ks, n, perm_left, bs_left);
    }
  } else {
    bpe_scatter_uniform<T, K><<<gr, 256, 0, st>>>(perm, ks, psi, n, Nx, Ny, U, psi_level, z1, Z.sx, Z.sy, Z.sz, p1, P.sx, P.sy, P.sz, perm_out,
                                                  bs_done ? nullptr : (T*)bs_out);
  }
  OCB_CUDA(cudaGetLastError());
  return OCB_OK;
}

template <class T, class K>
int bpe_impl(const ocb_grid* g, const ocb_field* b, int method, const ocb_field* zstar, const ocb_field* psid, int64_t* perm_out, void* bs_out,
             cudaStream_t st) {
  const int Nx = g->N[0], Ny = g->N[1], Nz = g->N[2];
  const long long n = (long long)Nx * Ny * Nz;
  OCB_REQUIRE(n < (1ll << 32), "the sort payload is a 32-bit index: at most 2^32 - 1 cells");
  if (method == OCB_BPE_THREE_DIMENSIONAL_SORT && !g->dzc && !getenv("OCB_BPE_GENERAL_EPILOGUE"))
    return bpe_sort3d_uniform<T, K>(g, b, zstar, psid, perm_out, bs_out, st);
  Scratch S(st);
  const int nblocks = (int)((n + RS_TILE - 1) / RS_TILE);
  T* bs = S.get<T>(n); T* zs_sorted = S.get<T>(n);
  K* ka = S.get<K>(n); K* kb = S.get<K>(n);
  uint32_t* ia = S.get<uint32_t>(n); uint32_t* ib = S.get<uint32_t>(n);
  uint32_t* table = S.get<uint32_t>(rs_table_words(nblocks));
  uint32_t* tot = S.get<uint32_t>((256ull * nblocks + SC_CHUNK - 1) / SC_CHUNK + 1 + 8 * 256);   // + the digit histograms of the onesweep path
  double* dV = S.get<double>(n); double* cum = S.get<double>(n); double* bdz = S.get<double>(n); double* psi = S.get<double>(n);
  double* dtot = S.get<double>((n + SC_CHUNK - 1) / SC_CHUNK + 1);
  double *den = nullptr, *nol = nullptr;
  if (method == OCB_BPE_HEAVISIDE_INTEGRAL) { den = S.get<double>(n); nol = S.get<double>(n); }
  if (!bs || !zs_sorted || !ka || !kb || !ia || !ib || !table || !tot || !dV || !cum || !bdz || !psi || !dtot ||
      (method == OCB_BPE_HEAVISIDE_INTEGRAL && (!den || !nol)))
    OCB_FAIL(OCB_ERR_ALLOC, "out of device memory for the reference-height scratch");
  const unsigned gr = (unsigned)((n + 255) / 256);
  // rank_by_buoyancy!: interior -> workspace -> sortperm
  HFld<T> B, Z, P;
  int rc = make_fld<T>(*b, &B, "b", false); if (rc != OCB_OK) return rc;
  rc = make_fld<T>(*zstar, &Z, "zstar", false); if (rc != OCB_OK) return rc;
  OCB_REQUIRE(B.p && Z.p, "b and zstar must be array fields");
  P.p = nullptr; P.sx = P.sy = P.sz = 0;
  if (psid && psid->data) { rc = make_fld<T>(*psid, &P, "psi_delta", false); if (rc != OCB_OK) return rc; }
  const T* b1 = B.p + (B.sx + B.sy + B.sz);        // element at index (1,1,1)
  K* ks; uint32_t* perm;       // the first pass makes the keys from the field, the last one writes the sorted values b*
  rc = radix_sort_pairs<K, FieldKeys<T, K>>(FieldKeys<T, K>{b1, B.sx, B.sy, B.sz, FastDiv((uint32_t)Nx), FastDiv((uint32_t)Ny)}, ka, ia, kb, ib, n, table, tot, nblocks, st, &ks, &perm,
                                            true, reinterpret_cast<K*>(bs));
  if (rc != OCB_OK) return rc;
  // volumes in sorted order and their cumulative sum (slot_centers!)
  const T* dzc1 = g->dzc ? (const T*)g->dzc + g->H[2] : nullptr;       // Δzᶜ at k = 1
  const T* zc1 = g->zc ? (const T*)g->zc + g->H[2] : nullptr;
  double volume = 0.0;     // A = ΣΔV / Lz ; on a RectilinearGrid ΣΔV = Lx Ly Lz exactly up to rounding of the products
  volume = g->L[0] * g->L[1] * g->L[2];
  const double A = volume / g->L[2];
  bpe_sorted_volumes<T><<<gr, 256, 0, st>>>(perm, n, Nx, Ny, (T)(g->d[0] * g->d[1]), dzc1, (T)g->d[2], dV);
  rc = device_scan<double, OP_ADD, false, false>(dV, cum, n, dtot, st); if (rc != OCB_OK) return rc;
  bpe_slot_centres<T><<<gr, 256, 0, st>>>(cum, dV, bs, n, g->z_bottom, A, method == OCB_BPE_THREE_DIMENSIONAL_SORT ? zs_sorted : nullptr, bdz);
  if (method == OCB_BPE_HEAVISIDE_INTEGRAL) {
    bpe_heaviside_prepare<T><<<gr, 256, 0, st>>>(bs, cum, dV, n, den, nol);
    rc = device_scan<double, OP_MAX, false, false>(den, den, n, dtot, st); if (rc != OCB_OK) return rc;
    rc = device_scan<double, OP_MIN, true, false>(nol, nol, n, dtot, st); if (rc != OCB_OK) return rc;
    bpe_heaviside_centres<T><<<gr, 256, 0, st>>>(den, nol, n, g->z_bottom, A, zs_sorted);
  }
  // Ψ at the slot faces: Ψface[m] = Σ_{q<m} b*[q] dz[q]  (fill_reference_potential!)
  rc = device_scan<double, OP_ADD, false, false>(bdz, psi, n, dtot, st); if (rc != OCB_OK) return rc;
  const T* z1 = Z.p + (Z.sx + Z.sy + Z.sz);
  T* p1 = P.p ? (T*)P.p + (P.sx + P.sy + P.sz) : nullptr;
  double* psi_level = S.get<double>(Nz);
  if (!psi_level) OCB_FAIL(OCB_ERR_ALLOC, "out of device memory for the reference-height scratch");
  if (p1) bpe_psi_levels<T><<<(Nz + 127) / 128, 128, 0, st>>>(cum, psi, bs, n, Nz, g->z_bottom, A, zc1, (T)g->z_bottom, (T)g->d[2], psi_level);
  if (method == OCB_BPE_THREE_DIMENSIONAL_SORT)
    bpe_scatter<T, true><<<gr, 256, 0, st>>>(perm, zs_sorted, cum, psi, bs, n, Nx, Ny, g->z_bottom, A, psi_level, (T*)z1, Z.sx, Z.sy, Z.sz,
                                            p1, P.sx, P.sy, P.sz);
  else
    bpe_scatter<T, false><<<gr, 256, 0, st>>>(perm, zs_sorted, cum, psi, bs, n, Nx, Ny, g->z_bottom, A, psi_level, (T*)z1, Z.sx, Z.sy, Z.sz,
                                             p1, P.sx, P.sy, P.sz);
  if (perm_out) bpe_perm_to_int64<<<gr, 256, 0, st>>>(perm, n, perm_out);
  if (bs_out) OCB_CUDA(cudaMemcpyAsync(bs_out, bs, n * sizeof(T), cudaMemcpyDeviceToDevice, st));
  OCB_CUDA(cudaGetLastError());
  return OCB_OK;
}

// ---- VerticalSort / ProfileLookup (scope table 8f) --------------------------------------------------------------------
// gather of a model-grid field into sorted order: dst[r] = field[perm[r]]  (s.source_height[perm], BPE :584-597)
template <class T>
__global__ void bpe_gather_perm(const T* __restrict__ f1, long long s0, long long s1, long long s2, int Nx, int Ny, const int64_t* __restrict__ perm,
                                long long n, T* __restrict__ dst) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;   // f1 pre-offset to index (1,1,1)
  if (t >= n) return;
  const long long c = perm[t] - 1;                                         // sortperm ranks are 1-based
  const int i = (int)(c % Nx), j = (int)((c / Nx) % Ny);
  const long long k = c / ((long long)Nx * Ny);
  dst[t] = f1[i * s0 + j * s1 + k * s2];
}

// profile_slot_faces (:497-507): midpoints between neighbouring heights, closed by the bottom and top of the domain;
// and the Ψ increments b✶[m] (faces[m+1] - faces[m]) of reference_potential_function (:359-368)
template <class T>
__global__ void bpe_profile_faces(const T* __restrict__ z, long long M, T z_bot, T z_top, T* __restrict__ faces) {
  const long long m = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (m > M) return;
  faces[m] = m == 0 ? z_bot : (m == M ? z_top : (z[m - 1] + z[m]) / T(2));
}
template <class T>
__global__ void bpe_profile_increments(const T* __restrict__ bs, const T* __restrict__ faces, long long M, double* __restrict__ inc) {
  const long long m = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (m < M) inc[m] = (double)bs[m] * ((double)faces[m + 1] - (double)faces[m]);
}
// 1-based searchsortedfirst / searchsortedlast over a non-decreasing array a[0..M)
template <class T> __device__ __forceinline__ long long ss_first(const T* a, long long M, T x) {   // first idx with a[idx] >= x, M+1 if none
  long long lo = 0, hi = M;
  while (lo < hi) { const long long mid = (lo + hi) >> 1; if (a[mid] < x) lo = mid + 1; else hi = mid; }
  return lo + 1;
}
template <class T> __device__ __forceinline__ long long ss_last(const T* a, long long M, T x) {    // last idx with a[idx] <= x, 0 if none
  long long lo = 0, hi = M;
  while (lo < hi) { const long long mid = (lo + hi) >> 1; if (a[mid] <= x) lo = mid + 1; else hi = mid; }
  return lo;
}
// assign_reference_height!(::ProfileLookup) (:550-582): every cell takes the mid-height of the run of slots of its
// buoyancy class, and Ψδ = Ψ(z) - Ψ(z✶) through the profile's own piecewise-linear Ψ
template <class T>
__global__ void bpe_profile_lookup(const T* __restrict__ b1, long long bs0, long long bs1, long long bs2, int Nx, int Ny, long long n,
                                   const T* __restrict__ bs, const T* __restrict__ faces, const double* __restrict__ psi_face, long long M,
                                   const T* __restrict__ zc1, T z_bot, T dz, T* __restrict__ z1, long long zs0, long long zs1, long long zs2,
                                   T* __restrict__ p1, long long ps0, long long ps1, long long ps2) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= n) return;
  const int i = (int)(t % Nx), j = (int)((t / Nx) % Ny);
  const long long k = t / ((long long)Nx * Ny);
  const T b = b1[i * bs0 + j * bs1 + k * bs2];
  long long lo = ss_first(bs, M, b);
  lo = lo < 1 ? 1 : (lo > M ? M : lo);
  const long long lol = lo - 1 < 1 ? 1 : lo - 1;
  const T dl = bs[lol - 1] - b, dh = bs[lo - 1] - b;
  const T bclass = (dl < 0 ? -dl : dl) < (dh < 0 ? -dh : dh) ? bs[lol - 1] : bs[lo - 1];
  const long long rf = ss_first(bs, M, bclass), rl = ss_last(bs, M, bclass);
  const T zstar = (faces[rf - 1] + faces[rl]) / T(2);                      // faces[run_first], faces[run_last + 1] (1-based)
  z1[i * zs0 + j * zs1 + k * zs2] = zstar;
  if (p1) {
    auto psi = [&](T zeta) {
      long long kk = ss_last(faces, M + 1, zeta);
      kk = kk < 1 ? 1 : (kk > M ? M : kk);
      return psi_face[kk - 1] + (double)bs[kk - 1] * ((double)zeta - (double)faces[kk - 1]);
    };
    const T zsrc = zc1 ? zc1[k] : (T)(z_bot + (T(k) + T(0.5)) * dz);
    p1[i * ps0 + j * ps1 + k * ps2] = (T)(psi(zsrc) - psi(zstar));
  }
}

template <class T>
int gather_impl(const ocb_field* src, const int64_t* perm, T* dst, long long n, cudaStream_t st) {
  HFld<T> F;
  int rc = make_fld<T>(*src, &F, "src", false);
  if (rc != OCB_OK) return rc;
  const int Nx = src->size[0] - 2 * src->halo[0], Ny = src->size[1] - 2 * src->halo[1], Nz = src->size[2] - 2 * src->halo[2];
  OCB_REQUIRE((long long)Nx * Ny * Nz == n, "ocb_gather_sorted: the permutation has %lld entries, the field %lld cells", n, (long long)Nx * Ny * Nz);
  const T* f1 = F.p + (F.sx + F.sy + F.sz);
  bpe_gather_perm<T><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(f1, F.sx, F.sy, F.sz, Nx, Ny, perm, n, dst);
  OCB_CUDA(cudaGetLastError());
  return OCB_OK;
}

template <class T>
int lookup_impl(const ocb_grid* g, const ocb_field* b, const T* bs, const T* zprof, long long M, const ocb_field* zstar, const ocb_field* psid,
                cudaStream_t st) {
  Scratch S(st);
  const int Nx = g->N[0], Ny = g->N[1], Nz = g->N[2];
  const long long n = (long long)Nx * Ny * Nz;
  HFld<T> B, Z, P;
  int rc = make_fld<T>(*b, &B, "b", false); if (rc != OCB_OK) return rc;
  rc = make_fld<T>(*zstar, &Z, "zstar", false); if (rc != OCB_OK) return rc;
  P.p = nullptr;
  if (psid && psid->data) { rc = make_fld<T>(*psid, &P, "psi_delta", false); if (rc != OCB_OK) return rc; }
  T* faces = S.get<T>(M + 1);
  double* inc = S.get<double>(M + 1);
  double* psi_face = S.get<double>(M + 1);
  double* tot = S.get<double>((M + SC_CHUNK - 1) / SC_CHUNK + 1);
  if (!faces || !inc || !psi_face || !tot) OCB_FAIL(OCB_ERR_ALLOC, "out of device memory for the profile-lookup scratch");
  const unsigned gm = (unsigned)((M + 1 + 255) / 256), gn = (unsigned)((n + 255) / 256);
  bpe_profile_faces<T><<<gm, 256, 0, st>>>(zprof, M, (T)g->z_bottom, (T)(g->z_bottom + g->L[2]), faces);
  bpe_profile_increments<T><<<gm, 256, 0, st>>>(bs, faces, M, inc);
  OCB_CUDA(cudaMemsetAsync(psi_face, 0, sizeof(double), st));
  rc = device_scan<double, OP_ADD, false, false>(inc, psi_face + 1, M, tot, st);     // Ψ at the faces: inclusive sums behind a leading 0
  if (rc != OCB_OK) return rc;
  const T* zc1 = g->zc ? (const T*)g->zc + g->H[2] : nullptr;
  bpe_profile_lookup<T><<<gn, 256, 0, st>>>(B.p + (B.sx + B.sy + B.sz), B.sx, B.sy, B.sz, Nx, Ny, n, bs, faces, psi_face, M, zc1, (T)g->z_bottom,
                                            (T)g->d[2], (T*)Z.p + (Z.sx + Z.sy + Z.sz), Z.sx, Z.sy, Z.sz,
                                            P.p ? (T*)P.p + (P.sx + P.sy + P.sz) : nullptr, P.sx, P.sy, P.sz);
  OCB_CUDA(cudaGetLastError());
  return OCB_OK;
}

}  // namespace

extern "C" int ocb_gather_sorted(int32_t dtype, const ocb_field* src, const int64_t* perm, void* dst, int64_t n, void* stream) {
  OCB_REQUIRE(src && perm && dst && n >= 0, "ocb_gather_sorted: bad argument");
  OCB_REQUIRE(src->kind == OCB_FIELD_ARRAY && src->data, "ocb_gather_sorted: the source must be an array field");
  OCB_REQUIRE(dtype == OCB_F32 || dtype == OCB_F64, "bad dtype");
  int rc = ocb_require_device();
  if (rc != OCB_OK) return rc;
  OcbDeviceGuard device_guard(src->data);
  if (n == 0) return OCB_OK;
  return dtype == OCB_F64 ? gather_impl<double>(src, perm, (double*)dst, n, (cudaStream_t)stream)
                          : gather_impl<float>(src, perm, (float*)dst, n, (cudaStream_t)stream);
}

extern "C" int ocb_bpe_profile_lookup(const ocb_grid* grid, const ocb_field* b, const void* b_profile, const void* z_profile, int64_t m,
                                      const ocb_field* zstar, const ocb_field* psi_delta, void* stream) {
  OCB_REQUIRE(grid && b && b_profile && z_profile && zstar, "ocb_bpe_profile_lookup: null argument");
  OCB_REQUIRE(m >= 1, "`ProfileLookup` needs a profile with at least one slot");
  OCB_REQUIRE(b->kind == OCB_FIELD_ARRAY && b->data && zstar->kind == OCB_FIELD_ARRAY && zstar->data, "b and zstar must be array fields");
  for (int d = 0; d < 3; ++d) OCB_REQUIRE(b->loc[d] == OCB_CENTER && zstar->loc[d] == OCB_CENTER, "b and zstar must be (Center, Center, Center) fields");
  int rc = ocb_require_device();
  if (rc != OCB_OK) return rc;
  OcbDeviceGuard device_guard(b->data);
  ocb_keep_scratch_pool();
  cudaStream_t st = (cudaStream_t)stream;
  return grid->dtype == OCB_F64 ? lookup_impl<double>(grid, b, (const double*)b_profile, (const double*)z_profile, m, zstar, psi_delta, st)
                                : lookup_impl<float>(grid, b, (const float*)b_profile, (const float*)z_profile, m, zstar, psi_delta, st);
}

extern "C" int ocb_sort_pairs(int32_t dtype, const void* keys_in, void* keys_out, uint32_t* perm_out, int64_t n, void* stream) {
  OCB_REQUIRE(keys_in && n >= 0 && n < (1ll << 32), "ocb_sort_pairs: bad argument");
  OCB_REQUIRE(dtype == OCB_F32 || dtype == OCB_F64, "bad dtype");
  int rc = ocb_require_device();
  if (rc != OCB_OK) return rc;
  OcbDeviceGuard device_guard(keys_in);
  if (n == 0) return OCB_OK;
  ocb_keep_scratch_pool();
  return dtype == OCB_F64 ? sort_impl<double, uint64_t>((const double*)keys_in, (double*)keys_out, perm_out, n, (cudaStream_t)stream)
                          : sort_impl<float, uint32_t>((const float*)keys_in, (float*)keys_out, perm_out, n, (cudaStream_t)stream);
}

extern "C" int ocb_bpe_reference_height(const ocb_grid* grid, const ocb_field* b, int32_t method, const ocb_field* zstar, const ocb_field* psi_delta,
                                        int64_t* perm_out, void* b_sorted_out, void* stream) {
  OCB_REQUIRE(grid && b && zstar, "ocb_bpe_reference_height: null argument");
  OCB_REQUIRE(method == OCB_BPE_THREE_DIMENSIONAL_SORT || method == OCB_BPE_HEAVISIDE_INTEGRAL, "unknown reference-height method %d", method);
  OCB_REQUIRE(b->kind == OCB_FIELD_ARRAY && b->data && zstar->kind == OCB_FIELD_ARRAY && zstar->data, "b and zstar must be array fields");
  for (int d = 0; d < 3; ++d) OCB_REQUIRE(b->loc[d] == OCB_CENTER && zstar->loc[d] == OCB_CENTER, "b and zstar must be (Center, Center, Center) fields");
  if (method == OCB_BPE_THREE_DIMENSIONAL_SORT)   // src/BackgroundPotentialEnergyEquation.jl:770-788
    OCB_REQUIRE(!grid->dzc, "`ThreeDimensionalSort` needs a grid with uniform cell volumes (regular spacing in every direction), but this grid is "
                            "stretched. Only `HeavisideIntegral()` supports stretched grids.");
  int rc = ocb_require_device();
  if (rc != OCB_OK) return rc;
  OcbDeviceGuard device_guard(b->data);
  ocb_keep_scratch_pool();
  cudaStream_t st = (cudaStream_t)stream;
  return grid->dtype == OCB_F64 ? bpe_impl<double, uint64_t>(grid, b, method, zstar, psi_delta, perm_out, b_sorted_out, st)
                                : bpe_impl<float, uint32_t>(grid, b, method, zstar, psi_delta, perm_out, b_sorted_out, st);
}
