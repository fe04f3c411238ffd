// K1c: the velocity-gradient kernel -- closure-dependent dissipation and the strain / rotation moduli.
//
// One sweep over u, v, w (and the closure's eddy-viscosity fields) emits any subset of
//   EPS   viscous_dissipation_rate_ccc            (src/KineticEnergyEquation.jl:427-453)   ν constant and/or ν_e fields
//   EISO  isotropic_viscous_dissipation_rate_ccc  (src/KineticEnergyEquation.jl:497-510)
//   SMOD  strain_rate_tensor_modulus_ccc          (src/FlowDiagnostics.jl:429-441)
//   OMOD  vorticity_tensor_modulus_ccc            (src/FlowDiagnostics.jl:575-587)
//   Q     Q_velocity_gradient_tensor_invariant_ccc (src/FlowDiagnostics.jl:708-712)
// as output fields and/or fused volume integrals, in Float32 or Float64, on a regular or stretched z axis
// (BASELINE.json configs[3]: 768x768x384 Float32, stretched z, SmagorinskyLilly / AMD-like ν_e).
//
// Same decomposition as the budget kernel (ocb_sweep_fused.cu): a block owns a 32 x BR column of cells and marches
// in z, a thread owns RY consecutive rows of one x column, every edge quantity (ffc / fcf / cff) is formed once by
// the cell that owns it and shared by shuffle (x), register or exchange row (y) and the register pipeline (z), tiles
// overlap by one cell in x and y.  Float32 parent rows are not 16-byte multiples in general (Nx + 2H = 774 floats in
// config 4), so the plane tiles cannot come by TMA: they are brought into the shared-memory ring by per-element
// cp.async (LDGSTS) copies issued NS planes ahead, committed per plane and awaited with cp.async.wait_group.
// Round 2 (TMAL = true): the tiles come by TMA after all.  Float32 with even parent row and column counts: two parent rows ARE a
// 16-byte multiple, so the tensor maps describe each plane as rows of 2 (Nx + 2H) elements and a tile is two boxes per field (the
// rows of one parity, the rows of the other; ocb_budget_sums.cuh has the same form).  Float64 rows of an even length are 16-byte
// multiples themselves: one box.  One thread issues the loads, nobody copies: config 4 2.48 -> 1.61 ms, its Float64 twin 4.35 -> 2.72 ms.
#include "ocb_sweep_strain.cuh"
#include "ocb_tma.cuh"
#include <new>

namespace {

constexpr int SD_EPS = 0, SD_EISO = 1, SD_SMOD = 2, SD_OMOD = 3, SD_Q = 4, SD_N = 5;
constexpr int FAM_E = 1, FAM_P = 2, FAM_M = 4;     // edge families: ν-weighted strain², (a+b)², (a-b)²
constexpr int STW = 36;                            // tile row pitch: i0-1 .. i0+32, plus one column either side for the pair alignment
constexpr int ZC = 32;                             // planes per staged window of z metrics
constexpr int STPAD = 4;                           // spare elements after each tile: the target of a thread's surplus copies
constexpr int NFMAX = 3 + OCB_MAX_CLOSURE_FIELDS;


template <class T>
struct StrainParams {
  int Nx, Ny, Nz, hx, hy, hz;
  int tiles_x, tiles_y, nchunks, kchunk, kz_lo, kz_hi, jy_lo, jy_hi;
  int cx, cy, cz;                        // last parent index every operand has along x / y / z (loads are clamped to it)
  int sy;                                // parent row pitch (elements), the same for every operand
  long long sz;                          // parent plane pitch
  T rdx, rdy, nu_c;
  const T *dzc, *dzf, *rdzc, *rdzf;      // pre-offset z tables: table[k] with the 1-based k
  const T* f[NFMAX];                     // u, v, w, ν_e fields (parent base pointers)
  T* out[SD_N];
  int slot[SD_N];
  long long os[3];
  double dxdy;
  double* partial;
  int nslots;
};

__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int B>
__device__ __forceinline__ void cp_async(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(dst), "l"(src), "n"(B) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ float t_sqrt(float x) { return sqrtf(x); }
__device__ __forceinline__ double t_sqrt(double x) { return sqrt(x); }

struct StrainMaps { CUtensorMap m[NFMAX]; };
constexpr int STWT = 40;                           // tile row pitch of the TMA form: 34 columns + up to 3 of box alignment, 16-byte rows
constexpr int strain_half_elems(int tr) { return ((STWT * (tr / 2) * 4 + 127) / 128) * 128 / 4; }   // one box, 128-byte aligned
constexpr int strain_tile_f64(int tr) { return ((STW * tr * 8 + 127) / 128) * 128 / 8; }               // Float64: one box of STW x tr

template <class T, int FAM, int NNU, int NW, int RY, int NS, int MINB, bool TMAL = false>
__global__ void __launch_bounds__(NW * 32, MINB)
ocb_strain_kernel(const __grid_constant__ StrainParams<T> P, const __grid_constant__ StrainMaps M) {
  constexpr bool F_E = FAM & FAM_E, F_P = FAM & FAM_P, F_M = FAM & FAM_M;
  constexpr int NFAM = (F_E ? 1 : 0) + (F_P ? 1 : 0) + (F_M ? 1 : 0);
  constexpr int IE = 0, IP = F_E ? 1 : 0, IM = IP + (F_P ? 1 : 0);       // family index -> slot in the small arrays
  constexpr int NF = 3 + NNU, BR = NW * RY, TR = BR + 2, NT = NW * 32;
  constexpr bool F32T = TMAL && sizeof(T) == 4;                // two boxes per tile; Float64 rows are 16-byte multiples: one box
  constexpr int HALF = strain_half_elems(TR);
  constexpr int TILE = F32T ? 2 * HALF : (TMAL ? strain_tile_f64(TR) : STW * TR + STPAD);
  static_assert(!F32T || TR % 2 == 0, "the Float32 TMA form takes two boxes of TR / 2 rows");
  constexpr int NXS = 2 * NFAM;                                          // exchanged per column: X12 and q23 of each family
  constexpr int NLD = ((STW / 2) * TR + NT - 1) / NT;                    // copies (element pairs) per thread, field and plane

  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* tiles = reinterpret_cast<T*>(smem_raw);                             // [NS][NF][TILE]
  T* xch = tiles + NS * NF * TILE;                                       // [2][NW][NXS][32]
  double* red = reinterpret_cast<double*>(xch + 2 * NW * NXS * 32);      // [NW]
  T* zts = reinterpret_cast<T*>(red + NW);                               // [4][ZC]
  uint64_t* full = reinterpret_cast<uint64_t*>(zts + 4 * ZC);              // [NS] (TMAL)

  const int lane = threadIdx.x, warp = threadIdx.y;
  const int tid = warp * 32 + lane;
  int bid = blockIdx.x;
  const int tx = bid % P.tiles_x; bid /= P.tiles_x;
  const int ty = bid % P.tiles_y; bid /= P.tiles_y;
  const int chunk = bid;
  const int i0 = 1 + 31 * tx, j0 = P.jy_lo + (BR - 1) * ty;
  const int k0 = P.kz_lo + chunk * P.kchunk;
  const int k1 = min(k0 + P.kchunk - 1, P.kz_hi);
  const int n_first = k0 - 1, n_last = k1 + 1;
  const int i = i0 + lane;
  const int jr0 = warp * RY;

  // ---- loader: this thread's share of the tile (same positions for every field and plane) -------------------------
  // Rows are copied as aligned element pairs (8-byte cp.async in Float32, 16-byte in Float64): the tile starts at the
  // even parent column at or before i0-1, and xoff = 0/1 shifts every shared-memory column index.
  const int px_raw = i0 - 1 + P.hx - 1, py = j0 - 1 + P.hy - 1;
  const int xoff = px_raw & 1, px = px_raw - xoff;
  int goff[NLD];
  uint32_t sb[NLD];
  const uint32_t tiles_u32 = s_u32(tiles);
#pragma unroll
  for (int t = 0; t < NLD; ++t) {
    const int e = tid + t * NT;
    const int row = e / (STW / 2), col = 2 * (e - row * (STW / 2));
    const bool ok = e < (STW / 2) * TR;
    const int gx = min(px + col, P.cx - 1), gy = min(py + row, P.cy);
    goff[t] = ok ? gx + gy * P.sy : 0;
    // a thread with no pair left in the last round copies something valid into the tile's spare slot instead:
    // every copy is then unconditional (no divergent branch around LDGSTS)
    sb[t] = (uint32_t)((ok ? row * STW + col : STW * TR) * sizeof(T));
  }
  // TMA form: box b holds the tile rows b, b + 2, .. (one parity of parent rows), read from the merged-row view at column
  // (parity) * rowlen + px_raw rounded down to 16 bytes; bo[b] = the columns lost to the rounding
  int bx[2] = {0, 0}, by[2] = {0, 0}, bo[2] = {0, 0};
  if (F32T) {
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      const int xr = ((py + b) & 1) * P.sy + px_raw;
      bx[b] = xr & ~3; bo[b] = xr & 3; by[b] = (py + b) >> 1;
    }
  }
  auto issue = [&](int n) {
    const int s = (n - n_first) % NS;
    if (TMAL) {                                          // (called by thread 0 only)
      mbar_expect_tx(&full[s], (uint32_t)(NF * (F32T ? STWT : STW) * TR * sizeof(T)));
      const int pz = n + P.hz - 1;                       // (planes a field does not have arrive as zeros)
#pragma unroll
      for (int f = 0; f < NF; ++f) {
        T* dst = tiles + (s * NF + f) * TILE;
        if (F32T) {
          tma_load_3d(dst, &M.m[f], &full[s], bx[0], by[0], pz);
          tma_load_3d(dst + HALF, &M.m[f], &full[s], bx[1], by[1], pz);
        } else {
          tma_load_3d(dst, &M.m[f], &full[s], px, py, pz);
        }
      }
      return;
    }
    const long long pz = min(n + P.hz - 1, P.cz);
#pragma unroll
    for (int f = 0; f < NF; ++f) {
      const uint32_t dst = tiles_u32 + (uint32_t)((s * NF + f) * TILE * sizeof(T));
      const T* src = P.f[f] + pz * P.sz;
      asm volatile("" : "+l"(src));                      // keep the plane base in a register: one address add per copy
#pragma unroll
      for (int t = 0; t < NLD; ++t) cp_async<2 * sizeof(T)>(dst + sb[t], src + goff[t]);
    }
  };
  if (TMAL) {
    if (tid == 0) {
      for (int q = 0; q < NS; ++q) mbar_init(&full[q], 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) for (int n = n_first; n < n_first + NS && n <= n_last; ++n) issue(n);
  } else {
    for (int n = n_first; n < n_first + NS; ++n) {      // one group per plane, empty past the end so the counting stays uniform
      if (n <= n_last) issue(n);
      cp_commit();
    }
  }
  // tile offsets of this lane's cell in the thread's rows, the row to the south (ro[0]) and the row to the north (ro[RY + 1])
  int ro[RY + 2];
#pragma unroll
  for (int q = 0; q < RY + 2; ++q) {
    const int tr = jr0 + q;
    ro[q] = F32T ? (tr & 1) * HALF + (tr >> 1) * STWT + bo[tr & 1] + lane + 1 : tr * STW + lane + xoff + 1;
  }
  uint32_t phase = 0;

  const T rdx = P.rdx, rdy = P.rdy, nu_c = P.nu_c;
  const T Q4 = T(0.25), H2 = T(0.5);
  // ---- pipeline state (per owned row) ----------------------------------------------------------------------------
  T up[RY], vp[RY], wp[RY], dgp[RY], nucp[RY], XNp[RY], YNp[RY];
  T accp[RY][NFAM > 0 ? NFAM : 1], facep[RY][NFAM > 0 ? NFAM : 1];
#pragma unroll
  for (int r = 0; r < RY; ++r) {
    up[r] = vp[r] = wp[r] = dgp[r] = nucp[r] = XNp[r] = YNp[r] = T(0);
#pragma unroll
    for (int a = 0; a < NFAM; ++a) accp[r][a] = facep[r][a] = T(0);
  }
  double acc[SD_N];
#pragma unroll
  for (int d = 0; d < SD_N; ++d) acc[d] = 0.0;
  const unsigned FULLMASK = 0xffffffffu;
  const bool lane_out = lane < 31 && i <= P.Nx;
  bool wout[RY];
  long long orow[RY];
#pragma unroll
  for (int r = 0; r < RY; ++r) {
    const int jr = jr0 + r, j = j0 + jr;
    wout[r] = lane_out && jr < BR - 1 && j <= P.jy_hi;
    orow[r] = i * P.os[0] + j * P.os[1];
  }

  const bool do_eiso = P.out[SD_EISO] || P.slot[SD_EISO] >= 0, do_smod = P.out[SD_SMOD] || P.slot[SD_SMOD] >= 0,
             do_omod = P.out[SD_OMOD] || P.slot[SD_OMOD] >= 0, do_q = P.out[SD_Q] || P.slot[SD_Q] >= 0;
  const bool integ = P.nslots > 0;
  int s = 0, par = 0;
  // z metrics: a window of ZC planes is staged in shared memory by the first warp (a global load consumed right
  // after the barrier would stall every warp of the block at once, once per plane)
#pragma unroll 2
  for (int n = n_first; n <= n_last; ++n) {
    const int zi = (n - n_first) % ZC;
    if (zi == 0 && tid < ZC) {
      const int q = min(n + tid, P.Nz + P.hz), km = max(q - 1, 1 - P.hz);       // the tables hold k = 1 - Hz .. Nz + Hz + 1
      zts[0 * ZC + tid] = __ldg(P.rdzf + q); zts[1 * ZC + tid] = __ldg(P.dzf + q);
      zts[2 * ZC + tid] = __ldg(P.rdzc + km); zts[3 * ZC + tid] = __ldg(P.dzc + km);
    }
    if (TMAL) {
      if (zi == 0) __syncthreads();                     // the z-metric window (uniform branch)
      mbar_wait(&full[s], phase);                       // plane n has landed
    } else {
      cp_wait<NS - 1>();                                // this thread's copies of plane n have landed ...
      __syncthreads();                                  // ... and everybody else's (and the z-metric window)
    }
    const T* Tt = tiles + s * (NF * TILE);
    const T* Tu = Tt, *Tv = Tt + TILE, *Tw = Tt + 2 * TILE;
    if (++s == NS) { s = 0; phase ^= 1; }
    const T rzf = zts[0 * ZC + zi], dzf_n = zts[1 * ZC + zi], rzc1 = zts[2 * ZC + zi], dzc1 = zts[3 * ZC + zi];
    const bool pv1 = (n - 1) >= k0 && (n - 1) <= k1;

    // ---- stage 1: edges owned by this thread's cells at plane n / z face n -------------------------------------------
    T uc[RY], vc[RY], wc[RY], dg[RY], nucc[RY], XN[RY], YN[RY];
    T X12[RY][NFAM > 0 ? NFAM : 1], Y13[RY][NFAM > 0 ? NFAM : 1], q23[RY][NFAM > 0 ? NFAM : 1];
    T nuc[RY], nuw[RY];                                  // eddy viscosity (summed over the closure's fields) at the cell and to its west
    // the thread's own column first: south / north neighbours inside the thread are then registers, not shared-memory reads
#pragma unroll
    for (int r = 0; r < RY; ++r) {
      const int o = ro[r + 1];
      uc[r] = Tu[o]; vc[r] = Tv[o]; wc[r] = Tw[o];
      nuc[r] = nuw[r] = T(0);
#pragma unroll
      for (int q = 0; q < NNU; ++q) { nuc[r] += Tt[(3 + q) * TILE + o]; nuw[r] += Tt[(3 + q) * TILE + o - 1]; }
    }
#pragma unroll
    for (int r = 0; r < RY; ++r) {
      const int oc = ro[r + 1];                           // this cell; oc - 1 west, oc + 1 east; ro[0] / ro[RY + 1]: the rows to the south / north
      const T ue = Tu[oc + 1], us = (r == 0) ? Tu[ro[0]] : uc[r > 0 ? r - 1 : 0];
      const T vw = Tv[oc - 1], vn = (r == RY - 1) ? Tv[ro[RY + 1]] : vc[r < RY - 1 ? r + 1 : r];
      const T ww = Tw[oc - 1], ws = (r == 0) ? Tw[ro[0]] : wc[r > 0 ? r - 1 : 0];
      const T a = (uc[r] - us) * rdy, b = (vc[r] - vw) * rdx;               // ∂y u, ∂x v   at ffc (i, j, n)
      const T c = (uc[r] - up[r]) * rzf, d = (wc[r] - ww) * rdx;            // ∂z u, ∂x w   at fcf (i, j, face n)
      const T e = (vc[r] - vp[r]) * rzf, f = (wc[r] - ws) * rdy;            // ∂z v, ∂y w   at cff
      const T p12 = (a + b) * (a + b), p13 = (c + d) * (c + d), p23 = (e + f) * (e + f);
      T q12[NFAM > 0 ? NFAM : 1], q13[NFAM > 0 ? NFAM : 1];
      nucc[r] = nu_c; XN[r] = YN[r] = T(0);
      T nffc = nu_c, nfcf = nu_c, ncff = nu_c;
      if (NNU > 0) {
        const T nc = nuc[r], nw_ = nuw[r];
        T ns = T(0), nsw = T(0);
        if (r == 0) {
#pragma unroll
          for (int q = 0; q < NNU; ++q) {
            const T* tn = Tt + (3 + q) * TILE;
            ns += tn[ro[0]]; nsw += tn[ro[0] - 1];
          }
        } else { ns = nuc[r > 0 ? r - 1 : 0]; nsw = nuw[r > 0 ? r - 1 : 0]; }
        nucc[r] = nu_c + nc;
        XN[r] = nw_ + nc; YN[r] = ns + nc;
        nffc = nu_c + Q4 * ((nsw + ns) + XN[r]);
        nfcf = nu_c + Q4 * (XNp[r] + XN[r]);
        ncff = nu_c + Q4 * (YNp[r] + YN[r]);
      }
      if (F_E) { q12[IE] = nffc * (Q4 * p12); q13[IE] = dzf_n * (nfcf * (Q4 * p13)); q23[r][IE] = dzf_n * (ncff * (Q4 * p23)); }
      if (F_P) { q12[IP] = p12; q13[IP] = p13; q23[r][IP] = p23; }
      if (F_M) { q12[IM] = (a - b) * (a - b); q13[IM] = (c - d) * (c - d); q23[r][IM] = (e - f) * (e - f); }
#pragma unroll
      for (int m = 0; m < NFAM; ++m) {
        X12[r][m] = q12[m] + __shfl_down_sync(FULLMASK, q12[m], 1);
        Y13[r][m] = q13[m] + __shfl_down_sync(FULLMASK, q13[m], 1);
      }
      const T s11 = (ue - uc[r]) * rdx, s22 = (vn - vc[r]) * rdy;
      dg[r] = s11 * s11 + s22 * s22;
    }

    // ---- y exchange between warps: row 0 of the warp to the north ------------------------------------------------------
    T* myx = xch + ((par * NW + warp) * NXS) * 32 + lane;
#pragma unroll
    for (int m = 0; m < NFAM; ++m) { myx[(2 * m) * 32] = X12[0][m]; myx[(2 * m + 1) * 32] = q23[0][m]; }
    __syncthreads();
    // every thread is done with raw stage s: refill it (plane n + NS)
    if (TMAL) {
      if (tid == 0 && n + NS <= n_last) issue(n + NS);
    } else {
      if (n + NS <= n_last) issue(n + NS);
      cp_commit();
    }
    T X12N[NFAM > 0 ? NFAM : 1], q23N[NFAM > 0 ? NFAM : 1];
#pragma unroll
    for (int m = 0; m < NFAM; ++m) { X12N[m] = T(0); q23N[m] = T(0); }
    if (warp + 1 < NW) {
      const T* nx_ = xch + ((par * NW + warp + 1) * NXS) * 32 + lane;
#pragma unroll
      for (int m = 0; m < NFAM; ++m) { X12N[m] = nx_[(2 * m) * 32]; q23N[m] = nx_[(2 * m + 1) * 32]; }
    }

    // ---- stage 2: finish the cells of plane n-1 -----------------------------------------------------------------------------
#pragma unroll
    for (int r = 0; r < RY; ++r) {
      const bool last = (r == RY - 1);
      T inpl[NFAM > 0 ? NFAM : 1], face[NFAM > 0 ? NFAM : 1], tot[NFAM > 0 ? NFAM : 1];
#pragma unroll
      for (int m = 0; m < NFAM; ++m) {
        const T x12n = last ? X12N[m] : X12[last ? r : r + 1][m];
        const T q23n = last ? q23N[m] : q23[last ? r : r + 1][m];
        inpl[m] = X12[r][m] + x12n;                          // the four ffc corners of cell (i, j, n)
        face[m] = Y13[r][m] + (q23[r][m] + q23n);            // the two fcf and two cff edges of z face n
        tot[m] = facep[r][m] + face[m];
      }
      const T s33 = (wc[r] - wp[r]) * rzc1;
      const T diag = dgp[r] + s33 * s33;
      const bool w1 = pv1 && wout[r];
      const long long o = orow[r] + (long long)(n - 1) * P.os[2];
      const double vol = (double)dzc1;
      T S2 = T(0), O2 = T(0);
      if (F_P) S2 = diag + H2 * (Q4 * accp[r][IP] + Q4 * tot[IP]);
      if (F_M) O2 = H2 * (Q4 * accp[r][IM] + Q4 * tot[IM]);
      // 2 ν_ccc Σ_aa² + 4 [mean_xy e12 + (mean_xz e13 + mean_yz e23) / Δz_c]   (the Δz_f weights are inside e13, e23)
      T vE = T(0), vI = T(0), vS = T(0), vO = T(0), vQ = T(0);
      if (F_E) vE = T(2) * nucp[r] * diag + (accp[r][IE] + rzc1 * tot[IE]);
      if (F_P && do_eiso) vI = T(2) * nucp[r] * S2;
      if (F_P && (do_smod || do_q)) vS = t_sqrt(S2);
      if (F_M && (do_omod || do_q)) vO = t_sqrt(O2);
      if (F_P && F_M && do_q) vQ = H2 * (vO * vO - vS * vS);              // through the roots, as the reference does
      // integrals: overlap cells and planes outside the chunk enter as exact zeros (a select, not a multiply by 0)
      if (integ) {                                        // (uniform) plans that only write fields skip the Float64 accumulation
        if (F_E) acc[SD_EPS] = fma((double)(w1 ? vE : T(0)), vol, acc[SD_EPS]);
        if (F_P && do_eiso) acc[SD_EISO] = fma((double)(w1 ? vI : T(0)), vol, acc[SD_EISO]);
        if (F_P && do_smod) acc[SD_SMOD] = fma((double)(w1 ? vS : T(0)), vol, acc[SD_SMOD]);
        if (F_M && do_omod) acc[SD_OMOD] = fma((double)(w1 ? vO : T(0)), vol, acc[SD_OMOD]);
        if (F_P && F_M && do_q) acc[SD_Q] = fma((double)(w1 ? vQ : T(0)), vol, acc[SD_Q]);
      }
      if (w1) {
        if (F_E && P.out[SD_EPS]) P.out[SD_EPS][o] = vE;
        if (F_P && P.out[SD_EISO]) P.out[SD_EISO][o] = vI;
        if (F_P && P.out[SD_SMOD]) P.out[SD_SMOD][o] = vS;
        if (F_M && P.out[SD_OMOD]) P.out[SD_OMOD][o] = vO;
        if (F_P && F_M && P.out[SD_Q]) P.out[SD_Q][o] = vQ;
      }
#pragma unroll
      for (int m = 0; m < NFAM; ++m) { accp[r][m] = inpl[m]; facep[r][m] = face[m]; }
      dgp[r] = dg[r]; nucp[r] = nucc[r]; XNp[r] = XN[r]; YNp[r] = YN[r];
      up[r] = uc[r]; vp[r] = vc[r]; wp[r] = wc[r];
    }
    par ^= 1;
  }
  if (!TMAL) cp_wait<0>();

  // ---- volume integrals: warp shuffle, then one partial per block and slot --------------------------------------------
  if (P.nslots > 0) {
#pragma unroll
    for (int d = 0; d < SD_N; ++d) {
      if (P.slot[d] < 0) continue;
      double v = acc[d];
      for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(FULLMASK, v, o);
      __syncthreads();
      if (lane == 0) red[warp] = v;
      __syncthreads();
      if (tid == 0) {
        double sum = 0.0;
        for (int w = 0; w < NW; ++w) sum += red[w];
        P.partial[(size_t)blockIdx.x * P.nslots + P.slot[d]] = sum * P.dxdy;
      }
    }
  }
}

int strain_diag_index(int diag) {
  switch (diag) {
    case OCB_DIAG_EPS: return SD_EPS;
    case OCB_DIAG_EPS_ISO: return SD_EISO;
    case OCB_DIAG_SMOD: return SD_SMOD;
    case OCB_DIAG_OMOD: return SD_OMOD;
    case OCB_DIAG_Q: return SD_Q;
    default: return -1;
  }
}

}  // namespace

struct ocb_strain_plan {
  int dtype;
  StrainParams<double> Pd;
  StrainParams<float> Pf;
  StrainMaps maps;                 // TMA form (Float32): one merged-row tensor map per operand
  bool tmal;
  int nw, ry, nblocks;
  size_t smem;
  void (*kd)(StrainParams<double>, StrainMaps);
  void (*kf)(StrainParams<float>, StrainMaps);
};

bool ocb_strain_handles_diag(int diag) { return strain_diag_index(diag) >= 0; }

template <class T> struct StrainShape;
template <> struct StrainShape<float> { static constexpr int NW = 4, RY = 3, NS = 3, MINB = 4; };
template <> struct StrainShape<double> { static constexpr int NW = 4, RY = 3, NS = 3, MINB = 2; };

template <class T> static void set_kernel(ocb_strain_plan* sp, void (*k)(StrainParams<T>, StrainMaps));
template <> void set_kernel<double>(ocb_strain_plan* sp, void (*k)(StrainParams<double>, StrainMaps)) { sp->kd = k; }
template <> void set_kernel<float>(ocb_strain_plan* sp, void (*k)(StrainParams<float>, StrainMaps)) { sp->kf = k; }

static void keep_params(ocb_strain_plan* sp, const StrainParams<double>& P) { sp->Pd = P; }
static void keep_params(ocb_strain_plan* sp, const StrainParams<float>& P) { sp->Pf = P; }

template <class T, int FAM, int NNU, int NW, int RY, int NS, int MINB>
static int strain_setup_shape(ocb_strain_plan* sp) {
  constexpr int NFAM = ((FAM & 1) ? 1 : 0) + ((FAM & 2) ? 1 : 0) + ((FAM & 4) ? 1 : 0);
  constexpr int NF = 3 + NNU, TR = NW * RY + 2, TILE = STW * TR + STPAD;
  sp->nw = NW; sp->ry = RY;
  if constexpr ((NW * RY) % 2 == 0 || sizeof(T) == 8) {
    if (sp->tmal) {                                       // TMA form: 128-byte-aligned boxes (two per field tile in Float32), and the ring's mbarriers
      constexpr int TILE_T = sizeof(T) == 4 ? 2 * strain_half_elems(TR) : strain_tile_f64(TR);
      sp->smem = sizeof(T) * ((size_t)NS * NF * TILE_T + 2 * NW * 2 * NFAM * 32 + 4 * ZC) + sizeof(double) * NW + sizeof(uint64_t) * NS;
      auto kt = ocb_strain_kernel<T, FAM, NNU, NW, RY, NS, MINB, true>;
      cudaFuncSetAttribute(kt, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
      OCB_CUDA(cudaFuncSetAttribute(kt, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sp->smem));
      set_kernel<T>(sp, kt);
      return OCB_OK;
    }
  }
  sp->tmal = false;
  sp->smem = sizeof(T) * ((size_t)NS * NF * TILE + 2 * NW * 2 * NFAM * 32 + 4 * ZC) + sizeof(double) * NW;
  auto k = ocb_strain_kernel<T, FAM, NNU, NW, RY, NS, MINB>;
  cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  OCB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sp->smem));
  set_kernel<T>(sp, k);
  return OCB_OK;
}

template <class T, int FAM, int NNU>
static int strain_setup(ocb_strain_plan* sp) {
  using S = StrainShape<T>;
  // Shapes measured on B200 (profiles/r1_summary.md): small blocks win -- 4 warps x 3 rows, four blocks per SM in
  // Float32 (2.45 ms for config 4; 8 x 3 x 2: 2.7 ms, 16 x 2 x 1: 3.9 ms), two in Float64 (4.3 ms; 8 x 2 x 1: 5.5 ms).
  if (const char* cfg = getenv("OCB_STRAIN_CONFIG")) {       // tuning runs: "NW,RY,MINB" for the config-4 kernel (eps + |S|, one nu_e field)
    if (FAM == 3 && NNU == 1) {
      int nw = 0, ry = 0, mb = 0, ns = 3;
      sscanf(cfg, "%d,%d,%d,%d", &nw, &ry, &mb, &ns);
#define OCB_SHAPE(A, B, C, D) if (nw == A && ry == B && mb == C && ns == D) return strain_setup_shape<T, 3, 1, A, B, D, C>(sp);
      OCB_SHAPE(4, 3, 4, 3) OCB_SHAPE(4, 3, 2, 3) OCB_SHAPE(8, 3, 2, 3) OCB_SHAPE(8, 2, 1, 3)
#undef OCB_SHAPE
    }
  }
  return strain_setup_shape<T, FAM, NNU, S::NW, S::RY, S::NS, S::MINB>(sp);
}

template <class T, int FAM>
static int strain_pick_nu(ocb_strain_plan* sp, int nnu) {
  switch (nnu) {
    case 0: return strain_setup<T, FAM, 0>(sp);
    case 1: return strain_setup<T, FAM, 1>(sp);
    default: return strain_setup<T, FAM, 2>(sp);
  }
}

static bool strain_operand_ok(const ocb_field& f, const ocb_grid* g) {
  if (f.kind != OCB_FIELD_ARRAY || !f.data || f.stride[0] != 1) return false;
  for (int d = 0; d < 3; ++d) if (f.loc[d] == OCB_NOTHING || f.halo[d] != g->H[d]) return false;
  return true;
}

template <class T>
static int strain_try_create(const ocb_grid* grid, const ocb_sweep_inputs* in, const ocb_request* reqs, int n,
                             const SweepParams<T>& SP, ocb_strain_plan** out) {
  *out = nullptr;
  if (getenv("OCB_DISABLE_FUSED") || getenv("OCB_DISABLE_STRAIN")) return OCB_OK;
  // ---- envelope --------------------------------------------------------------------------------------------------
  int fam = 0;
  int which_out[SD_N], which_slot[SD_N];
  for (int d = 0; d < SD_N; ++d) which_out[d] = which_slot[d] = -1;
  bool any_out = false;
  long long os[3] = {0, 0, 0};
  for (int r = 0; r < n; ++r) {
    const int d = strain_diag_index(reqs[r].diag);
    if (d < 0 || (reqs[r].flags & OCB_REQ_PERTURBATION)) return OCB_OK;
    const DReq& R = SP.req[r];
    if (R.lo[0] != 1 || R.hi[0] != grid->N[0]) return OCB_OK;                                                // full x extent
    for (int a = 1; a < 3; ++a) if (R.lo[a] != SP.req[0].lo[a] || R.hi[a] != SP.req[0].hi[a]) return OCB_OK;  // one y and z window
    if (R.out) {
      if (which_out[d] >= 0) return OCB_OK;
      which_out[d] = r;
      if (any_out && (R.os[0] != os[0] || R.os[1] != os[1] || R.os[2] != os[2])) return OCB_OK;               // one layout for all outputs
      any_out = true;
      for (int a = 0; a < 3; ++a) os[a] = R.os[a];
    }
    if (R.slot >= 0) { if (which_slot[d] >= 0) return OCB_OK; which_slot[d] = r; }
    fam |= (d == SD_EPS) ? FAM_E : (d == SD_OMOD) ? FAM_M : (d == SD_Q) ? (FAM_P | FAM_M) : FAM_P;
  }
  if (!strain_operand_ok(in->u, grid) || !strain_operand_ok(in->v, grid) || !strain_operand_ok(in->w, grid)) return OCB_OK;
  if (in->u.loc[0] != OCB_FACE || in->v.loc[1] != OCB_FACE || in->w.loc[2] != OCB_FACE) return OCB_OK;
  const bool needs_nu = fam & FAM_E || which_out[SD_EISO] >= 0 || which_slot[SD_EISO] >= 0;
  const int nnu = needs_nu ? in->closure.n_nu_fields : 0;
  for (int q = 0; q < nnu; ++q) {
    const ocb_field& f = in->closure.nu_fields[q];
    if (!strain_operand_ok(f, grid) || f.loc[0] != OCB_CENTER || f.loc[1] != OCB_CENTER || f.loc[2] != OCB_CENTER) return OCB_OK;
  }

  ocb_strain_plan* sp = new (std::nothrow) ocb_strain_plan();
  if (!sp) OCB_FAIL(OCB_ERR_ALLOC, "out of host memory");
  memset((void*)sp, 0, sizeof(*sp));
  sp->dtype = grid->dtype;
  // Float32 operands whose parent rows and row counts are even: the tiles come by TMA (two boxes per field over the merged-row view)
  sp->tmal = false;
  if (!getenv("OCB_STRAIN_NO_TMA") && get_encode_fn()) {
    auto compat = [&](const ocb_field& fl) { return sizeof(T) == 4 ? tma_compatible_f32(fl, grid) : tma_compatible_f64(fl, grid); };
    bool ok = compat(in->u) && compat(in->v) && compat(in->w);
    for (int q = 0; q < nnu; ++q) ok = ok && compat(in->closure.nu_fields[q]);
    sp->tmal = ok;
  }
  int rc;
  // families are compiled in the combinations the diagnostics ask for: {E}, {P}, {E,P} (config 4), {M}, {P,M}, {E,P,M}
  if (fam == FAM_E) rc = strain_pick_nu<T, FAM_E>(sp, nnu);
  else if (fam == FAM_P) rc = strain_pick_nu<T, FAM_P>(sp, nnu);
  else if (fam == (FAM_E | FAM_P)) rc = strain_pick_nu<T, FAM_E | FAM_P>(sp, nnu);
  else if (fam == FAM_M) rc = strain_setup<T, FAM_M, 0>(sp);
  else if (fam == (FAM_P | FAM_M)) rc = strain_pick_nu<T, FAM_P | FAM_M>(sp, nnu);
  else rc = strain_pick_nu<T, FAM_E | FAM_P | FAM_M>(sp, nnu);
  if (rc != OCB_OK) { delete sp; return rc; }

  StrainParams<T> P;
  memset((void*)&P, 0, sizeof(P));
  P.Nx = grid->N[0]; P.Ny = grid->N[1]; P.Nz = grid->N[2];
  P.hx = grid->H[0]; P.hy = grid->H[1]; P.hz = grid->H[2];
  const int BR = sp->nw * sp->ry;
  P.tiles_x = (P.Nx + 30) / 31;
  P.jy_lo = SP.req[0].lo[1]; P.jy_hi = SP.req[0].hi[1];
  P.tiles_y = ((P.jy_hi - P.jy_lo + 1) + BR - 2) / (BR - 1);
  const int tiles = P.tiles_x * P.tiles_y;
  P.kz_lo = SP.req[0].lo[2]; P.kz_hi = SP.req[0].hi[2];
  const int nz_win = P.kz_hi - P.kz_lo + 1;
  int nchunks = (ocb_sm_count() * 8 + tiles - 1) / tiles;
  if (nchunks < 1) nchunks = 1;
  int kchunk = (nz_win + nchunks - 1) / nchunks;
  if (kchunk < 32) kchunk = nz_win < 32 ? nz_win : 32;
  P.kchunk = kchunk;
  P.nchunks = (nz_win + kchunk - 1) / kchunk;
  sp->nblocks = tiles * P.nchunks;
  P.rdx = T(1.0 / grid->d[0]); P.rdy = T(1.0 / grid->d[1]); P.nu_c = T(in->closure.nu_const);
  P.dzc = SP.g.dzc; P.dzf = SP.g.dzf; P.rdzc = SP.g.rdzc; P.rdzf = SP.g.rdzf;
  P.dxdy = grid->d[0] * grid->d[1];
  P.nslots = SP.nslots;
  const ocb_field* f[NFMAX] = {&in->u, &in->v, &in->w, &in->closure.nu_fields[0], &in->closure.nu_fields[1]};
  P.cx = P.cy = P.cz = 1 << 30;
  P.sy = (int)f[0]->stride[1]; P.sz = f[0]->stride[2];
  bool pairs = P.sy % 2 == 0 && P.sz % 2 == 0;       // rows are copied as aligned element pairs
  for (int a = 0; a < 3 + nnu; ++a) {
    P.f[a] = (const T*)f[a]->data;
    // one row / plane pitch for every operand (periodic x and y: u, v, w and the ccc fields share a parent shape)
    if (f[a]->stride[1] != P.sy || f[a]->stride[2] != P.sz || f[a]->stride[1] > (1 << 30)) { delete sp; return OCB_OK; }
    if ((uintptr_t)f[a]->data % (2 * sizeof(T))) pairs = false;
    if (f[a]->size[0] - 1 < P.cx) P.cx = f[a]->size[0] - 1;
    if (f[a]->size[1] - 1 < P.cy) P.cy = f[a]->size[1] - 1;
    if (f[a]->size[2] - 1 < P.cz) P.cz = f[a]->size[2] - 1;
  }
  if ((long long)P.sy * (P.cy + 1) > (1LL << 31) - 64) { delete sp; return OCB_OK; }   // in-plane offsets are 32-bit
  if (P.cx % 2 == 0 || !pairs) { delete sp; return OCB_OK; }                            // odd row length / unaligned base: generic kernel
  for (int a = 0; a < 3; ++a) P.os[a] = os[a];
  for (int d = 0; d < SD_N; ++d) {
    P.out[d] = which_out[d] >= 0 ? (T*)SP.req[which_out[d]].out : nullptr;
    P.slot[d] = which_slot[d] >= 0 ? SP.req[which_slot[d]].slot : -1;
  }
  if (sp->tmal) {
    EncodeTiledFn enc = get_encode_fn();
    const int TR = sp->nw * sp->ry + 2;
    for (int a = 0; a < NFMAX; ++a) {
      const ocb_field* fa = a < 3 + nnu ? f[a] : f[0];
      rc = sizeof(T) == 4 ? make_map_f32(enc, *fa, TR, &sp->maps.m[a], STWT) : make_map_f64(enc, *fa, TR, &sp->maps.m[a], STW);
      if (rc != OCB_OK) { delete sp; return rc; }
    }
  }
  keep_params(sp, P);
  *out = sp;
  return OCB_OK;
}

int ocb_strain_try_create(const ocb_grid* grid, const ocb_sweep_inputs* in, const ocb_request* reqs, int n,
                          const SweepParams<double>& P, ocb_strain_plan** out) {
  return strain_try_create<double>(grid, in, reqs, n, P, out);
}
int ocb_strain_try_create(const ocb_grid* grid, const ocb_sweep_inputs* in, const ocb_request* reqs, int n,
                          const SweepParams<float>& P, ocb_strain_plan** out) {
  return strain_try_create<float>(grid, in, reqs, n, P, out);
}

int ocb_strain_nblocks(const ocb_strain_plan* sp) { return sp->nblocks; }

int ocb_strain_execute(ocb_strain_plan* sp, double* partial, cudaStream_t st) {
  dim3 block(32, sp->nw, 1), grid(sp->nblocks, 1, 1);
  if (sp->dtype == OCB_F64) {
    sp->Pd.partial = partial;
    sp->kd<<<grid, block, sp->smem, st>>>(sp->Pd, sp->maps);
  } else {
    sp->Pf.partial = partial;
    sp->kf<<<grid, block, sp->smem, st>>>(sp->Pf, sp->maps);
  }
  OCB_CUDA(cudaGetLastError());
  return OCB_OK;
}

void ocb_strain_destroy(ocb_strain_plan* sp) { delete sp; }
