// K1 fast path: the register-pipelined, TMA-fed budget kernel.
//
// One sweep over u, v, w, b, p emits any subset of the cell-centred budget terms
//   K   kinetic_energy_ccc                     (src/KineticEnergyEquation.jl:37-39)
//   ADV uᵢ∂ⱼuⱼuᵢᶜᶜᶜ, Centered 2nd order         (src/KineticEnergyEquation.jl:147-152)
//   PR  uᵢ∂ᵢpᶜᶜᶜ                               (src/KineticEnergyEquation.jl:304-309)
//   WB  uᵢbᵢᶜᶜᶜ (vertical gravity)             (src/KineticEnergyEquation.jl:362-367)
//   EPS viscous_dissipation_rate_ccc           (src/KineticEnergyEquation.jl:427-453)
//   CHI tracer_variance_dissipation_rate_ccc   (src/TracerVarianceEquation.jl:146-161)
//   CDF c∇_dot_qᶜ                              (src/TracerVarianceEquation.jl:93-98)
//   SP  shear_production_rate(_z)_ccc about horizontal-mean profiles U(z), V(z), W(z)
//                                              (src/TurbulentKineticEnergyEquation.jl:224-240, 285-289)
//   EPV ertel_potential_vorticity_fff          (src/FlowDiagnostics.jl:216-233), periodic x and y
// as output fields and/or fused volume integrals, for Float64 array operands on a regularly spaced grid with
// a constant-coefficient closure.  Everything else goes to the generic kernel (ocb_sweep.cu).
//
// Design (B200):
//  * a block owns a (32 x BR) column of cells and marches in z; each thread owns RY consecutive rows of one x
//    column, so y neighbours are mostly registers and x neighbours are neighbouring lanes;
//  * every input plane tile (36 x (BR+2) with the x/y halo ring) is brought into a shared-memory ring by TMA
//    (cp.async.bulk.tensor.3d straight from the Oceananigans parent arrays, halos included), NS planes ahead,
//    completion signalled on an mbarrier -- no thread issues a global load;
//  * each edge (ffc/fcf/cff) and face quantity is formed ONCE by the cell that owns it and shared: along x by
//    warp shuffle, along y by registers (inside a thread) or a double-buffered shared exchange row (between
//    warps), along z by the register pipeline.  That removes the 4x corner recomputation of the per-cell
//    reference formulation;
//  * tiles overlap by one cell in x and y (lane 31 and the block's last row only produce shared quantities),
//    so no block ever needs another block's intermediates;
//  * volume integrals: per-thread Float64 accumulators, warp shuffle + shared reduction, one partial per
//    block and slot, summed in a fixed order by ocb_finalize_partials_kernel (deterministic run to run).
#include "ocb_sweep_fused.cuh"
#include "ocb_tma.cuh"
#include <new>
#include <type_traits>

namespace {

constexpr int FD_K = 0, FD_ADV = 1, FD_PR = 2, FD_WB = 3, FD_EPS = 4, FD_CHI = 5, FD_CDF = 6, FD_SP = 7, FD_EPV = 8,
              FD_EPSP = 9, FD_TKE = 10, FD_N = 11;   // EPSP / TKE: class-sum kernel only (the dissipation of u - U next to that of u; TKE)
constexpr int FD_BUDGET = 0x7f;   // the seven KE / tracer-variance budget terms
constexpr int NXCH = 8;            // exchanged quantities per column: P12, P23 | X12, e23, Fv (OUT only) | XS23, o13, o12 (EPV)
constexpr int TW = 36;             // tile width: 32 lanes + west/east halo + 1 (the TMA box must start 16-byte aligned) -> 36

struct FusedOut { double* out; int slot; };

struct FusedParams {
  int Nx, Ny, Nz;
  int hx, hy, hz;                  // parent index of 1-based index i is i + hx - 1
  int tiles_x, tiles_y, nchunks, kchunk;
  int kz_lo, kz_hi;                // 1-based inclusive z window (the whole interior, or a z slab of it)
  int jy_lo, jy_hi;                // 1-based inclusive y window (rows next to a slab edge are swept separately)
  double rdx, rdy, rdz, nu, kappa, ghat_z, V;
  int rowlen;                      // Float32-input class sums: elements per parent row (the merged-row view's odd rows start there)
  double kappa_scale;              // class-sum NUF variant: kappa at a face = kappa + kappa_scale x mean of nu_e (kappa_e = nu_e / Pr)
  FusedOut o[FD_N];
  long long os[3];                 // element strides shared by every output field
  double* partial;
  int nslots;
  // horizontal-mean profiles U(z), V(z), W(z) (pre-offset: profile[k] with the 1-based k; nullptr = ZeroField) and
  // whether the dissipation request sees the perturbation velocities u - U (lazy `u - U`, src/Oceanostics.jl:157-173)
  const double *Uz, *Vz, *Wz;
  int eps_pert;
  double f[3];                     // Coriolis vector of the potential-vorticity term
  int epv_top;                     // 1: the fff window also holds the top face Nz+1 (Bounded z)
  // class-sum kernel only.  y_ring: the y window is a whole periodic extent / one slab of a periodic ring (interior item
  // weights on every row).  gate: this rank's halo flag words {from low, from high, timed out} -- the first and last tile
  // rows wait until they reach gate_step (ocb_plan_execute_ring); nullptr: no waiting.
  // stretched z (ocb_budget_kernel<.., STR = true>): pre-offset metric tables, table[k] with the 1-based k, k = 1 - Hz .. Nz + Hz
  const double *dzc, *dzf, *rdzc, *rdzf;
  int y_ring;
  long long* gate;
  long long gate_step, gate_timeout;
};

template <int NW, int RY>
struct Geo {
  static constexpr int BR = NW * RY;                 // rows a block works on (the last one is overlap only)
  static constexpr int TR = BR + 2;                  // tile rows with the south/north halo
  static constexpr int TILE = ((TW * TR * 8 + 127) / 128) * 128 / 8;   // doubles per field tile, 128 B aligned
};

// MASK bit d set <=> diagnostic d is requested.  NS = TMA ring depth.
// OUT = true : some request materialises a field -- every cell value is formed (plane n-1 when plane n arrives,
//              ADV one plane later) and stored; integrals accumulate the same values.
// OUT = false: integrals only -- the SAME per-face / per-edge / per-cell terms are accumulated directly, each weighted
//              by the number of this block's output cells it belongs to (Σ_cells ½(F_i + F_{i+1}) = Σ_faces m_i F_i):
//              a pure re-ordering of the summation that needs no neighbour's face value, so most shuffles, three of
//              the five exchanged rows and seven pipeline registers per row disappear.
// (A warp-specialised variant -- producer warp + per-pair mbarriers instead of the block barrier -- was measured
//  slower on B200, 55.9 vs 66.7 Gcells/s, and removed; see profiles/r1_summary.md.)
// STR = true (with OUT): z is a stretched axis.  What changes is where a z metric sits (Oceananigans' area / volume weighted
// forms, restated in oracle/kernels.py): face gradients take 1/dzf of their face, cell differences 1/dzc of their cell, the fcf / cff
// strain edges and the z-face tracer gradients enter a cell with weight dzf(face) / dzc(cell), the advecting transports of the w
// equation are dzc-weighted means (Ax = dy dzc), and integrals weigh a cell by dzc.  The metrics are uniform loads per plane.
// TIN = float: Float32 fields in and out (the arithmetic and the integrals stay Float64): tiles as two TMA boxes per field over the
// merged-row view, read through per-row offsets -- the form of ocb_budget_sums.cuh and ocb_sweep_strain.cu.
template <int MASK, int NW, int RY, int NS, bool OUT, bool STR = false, class TIN = double>
__global__ void __launch_bounds__(NW * 32, (RY == 1 ? (NW <= 8 ? 2 : 1) : (RY == 2 && NW <= 6 ? 2 : 1)))
ocb_budget_kernel(const __grid_constant__ FusedParams P, const __grid_constant__ CUtensorMap mu, const __grid_constant__ CUtensorMap mv,
                  const __grid_constant__ CUtensorMap mw, const __grid_constant__ CUtensorMap mb, const __grid_constant__ CUtensorMap mp,
                  const __grid_constant__ CUtensorMap /* nu_e: class-sum kernel only */) {
  using G = Geo<NW, RY>;
  constexpr bool H_K = MASK & (1 << FD_K), H_ADV = MASK & (1 << FD_ADV), H_PR = MASK & (1 << FD_PR), H_WB = MASK & (1 << FD_WB),
                 H_EPS = MASK & (1 << FD_EPS), H_CHI = MASK & (1 << FD_CHI), H_CDF = MASK & (1 << FD_CDF), H_SP = MASK & (1 << FD_SP),
                 H_EPV = MASK & (1 << FD_EPV);
  constexpr bool PZ = H_SP;                            // kernels with the shear-production term carry the mean profiles
  constexpr bool NEED_B = H_WB || H_CHI || H_CDF || H_EPV, NEED_P = H_PR;
  constexpr int NF = 3 + (NEED_B ? 1 : 0) + (NEED_P ? 1 : 0);
  constexpr int F_B = 3, F_P = NEED_B ? 4 : 3;
  constexpr int NX = H_EPV ? NXCH : 5;
  // The fff point a thread reports is the north-east-top corner (i+1, j+1, n) of its cell, so that all sharing runs
  // the same way as for the cell terms (east by shuffle, north by register / exchange row); to reach the points
  // i = 1 and j = jy_lo the tiling then starts one cell earlier, in the halo.
  constexpr int SH = H_EPV ? 1 : 0;
  constexpr bool F32IN = sizeof(TIN) == 4;
  constexpr int TWK = F32IN ? 40 : TW;
  constexpr int HALF = ((TWK * (G::TR / 2) * 4 + 127) / 128) * 128 / 4;      // Float32: one of the tile's two boxes, 128-byte aligned
  constexpr int TILE = F32IN ? 2 * HALF : G::TILE;
  static_assert(!F32IN || (G::TR % 2 == 0 && !H_EPV && OUT), "Float32: two boxes of TR / 2 rows, the seven budget terms, the cell-value form");

  extern __shared__ __align__(1024) double smem_d[];
  TIN* tiles = reinterpret_cast<TIN*>(smem_d);                                 // [NS][NF][TILE]
  double* xch = reinterpret_cast<double*>(tiles + NS * NF * TILE);             // [2][NW][NX][32]
  double* red = xch + 2 * NW * NX * 32;                                        // [NW]
  uint64_t* full = reinterpret_cast<uint64_t*>(red + NW);                      // [NS]

  const int lane = threadIdx.x, warp = threadIdx.y;
  const int tid = warp * 32 + lane;
  int bid = blockIdx.x;
  const int tx = bid % P.tiles_x; bid /= P.tiles_x;
  const int ty = bid % P.tiles_y; bid /= P.tiles_y;
  const int chunk = bid;
  const int i0 = 1 - SH + 31 * tx, j0 = P.jy_lo - SH + (G::BR - 1) * ty;
  const int k0 = P.kz_lo + chunk * P.kchunk;
  const int k1 = min(k0 + P.kchunk - 1, P.kz_hi);
  const int n_first = k0 - 1, n_last = k1 + 2;
  const int i = i0 + lane;
  const int jr0 = warp * RY;                           // first block-row of this thread

  // ---- TMA ring ------------------------------------------------------------------------------------------
  constexpr uint32_t STAGE_BYTES = (uint32_t)(NF * TWK * G::TR * sizeof(TIN));
  // parent coordinates of the tile's first element; the TMA box start must be 16-byte aligned in x, so the box
  // starts at the even element at or before it and every shared-memory column index is shifted by xoff
  const int px_raw = i0 - 1 + P.hx - 1, py = j0 - 1 + P.hy - 1;
  const int xoff = px_raw & 1, px = px_raw - xoff;
  int bx[2] = {0, 0}, by[2] = {0, 0}, bo[2] = {0, 0};      // Float32: the two boxes' start (merged-row view) and the columns lost to its alignment
  if (F32IN) {
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      const int xr = ((py + b) & 1) * P.rowlen + px_raw;
      bx[b] = xr & ~3; bo[b] = xr & 3; by[b] = (py + b) >> 1;
    }
  }
  auto issue = [&](int n) {
    const int s = (n - n_first) % NS;
    TIN* dst = tiles + s * (NF * TILE);
    mbar_expect_tx(&full[s], STAGE_BYTES);
    const int pz = n + P.hz - 1;
#pragma unroll
    for (int b = 0; b < (F32IN ? 2 : 1); ++b) {
      TIN* d = dst + b * HALF;
      const int x = F32IN ? bx[b] : px, y = F32IN ? by[b] : py;
      tma_load_3d(d + 0 * TILE, &mu, &full[s], x, y, pz);
      tma_load_3d(d + 1 * TILE, &mv, &full[s], x, y, pz);
      tma_load_3d(d + 2 * TILE, &mw, &full[s], x, y, pz);
      if (NEED_B) tma_load_3d(d + F_B * TILE, &mb, &full[s], x, y, pz);
      if (NEED_P) tma_load_3d(d + F_P * TILE, &mp, &full[s], x, y, pz);
    }
  };
  // tile offsets of the column WEST of this lane's cell in the rows jr0 .. jr0 + RY + 1 (south row, the thread's rows, north row):
  // p = T + ro[q] gives p[0] west, p[1] centre, p[2] east
  int ro[RY + 2];
#pragma unroll
  for (int q = 0; q < RY + 2; ++q) {
    const int tr = jr0 + q;
    ro[q] = F32IN ? (tr & 1) * HALF + (tr >> 1) * TWK + bo[tr & 1] + lane : tr * TW + lane + xoff;
  }
  if (tid == 0) {
    for (int s = 0; s < NS; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) {
    for (int n = n_first; n < n_first + NS && n <= n_last; ++n) issue(n);
  }

  // metric and closure constants, folded so that every term below is a bare sum of products; the remaining
  // per-diagnostic scale factor (1/4 for K, 1/8 for ADV, ...) is applied to stored values and to the block partial
  const double rdx = P.rdx, rdy = P.rdy, rdz = P.rdz;
  const double rdx2 = rdx * rdx, rdy2 = rdy * rdy, rdz2 = rdz * rdz;
  const double hsn = 0.5 * sqrt(P.nu);                       // strain scale: (hsn * s)^2 = nu/4 * s^2
  const double ex = rdx * hsn, ey = rdy * hsn, ez = rdz * hsn;
  const double dx2 = 2.0 * P.nu * rdx2, dy2 = 2.0 * P.nu * rdy2, dz2 = 2.0 * P.nu * rdz2;   // diagonal strain weights
  const double SC_K = 0.25, SC_ADV = 0.125, SC_PR = 0.5, SC_WB = 0.25 * P.ghat_z, SC_EPS = 1.0, SC_CHI = P.kappa,
               SC_CDF = 2.0 * P.kappa;
  // ---- pipeline state (per owned row) -----------------------------------------------------------------------
  double up[RY], vp[RY], wp[RY], bp[RY], pp[RY];        // raw, previous plane
  double P13p[RY], P23p[RY];                            // fcf / cff momentum-flux products at the previous face
  double dU[RY], dV[RY], dW[RY];                        // in-plane parts of the flux divergences (previous plane)
  double Wwpp[RY], Apart[RY], Fvp[RY];                  // ADV pipeline (Apart, Fvp: OUT only)
  double Eacc[RY], Kacc[RY], PRacc[RY], CHacc[RY], CDacc[RY], hzp[RY];   // all but CDacc: OUT only
  double ucp[RY], vcp[RY];                              // centred perturbation velocities of the previous plane (SP)
  double aS_c[RY];
  double Dxp[RY], Dyp[RY], Sbp[RY], NE12p[RY];          // EPV: b-box sums and the corner ffc vorticity of the previous plane
  double Efp[RY], Gzp[RY], Gcp[RY];                     // STR: dzf-weighted edge sums / tracer z gradients of the previous face
#pragma unroll
  for (int r = 0; r < RY; ++r) {
    Efp[r] = Gzp[r] = Gcp[r] = 0.0;
    up[r] = vp[r] = wp[r] = bp[r] = pp[r] = 0.0;
    P13p[r] = P23p[r] = dU[r] = dV[r] = dW[r] = Wwpp[r] = Apart[r] = Fvp[r] = 0.0;
    Eacc[r] = Kacc[r] = PRacc[r] = CHacc[r] = CDacc[r] = hzp[r] = 0.0;
    ucp[r] = vcp[r] = aS_c[r] = 0.0;
    Dxp[r] = Dyp[r] = Sbp[r] = NE12p[r] = 0.0;
  }
  double acc[FD_N];
#pragma unroll
  for (int d = 0; d < FD_N; ++d) acc[d] = 0.0;
  // integral-only mode: unweighted per-row sums by multiplicity class (x faces / fcf edges: mx, y faces / cff edges: my,
  // ffc edges: mxy, cell and z-face terms: wrow); the class weights are applied once, after the plane loop
  double aE_xy[RY], aE_x[RY], aE_y[RY], aE_c[RY], aK_x[RY], aK_y[RY], aK_z[RY], aP_x[RY], aP_y[RY], aP_z[RY];
  double aC_x[RY], aC_y[RY], aC_z[RY], aW_z[RY], aD_c[RY], aA_x[RY], aA_y[RY], aA_z[RY];
#pragma unroll
  for (int r = 0; r < RY; ++r) {
    aE_xy[r] = aE_x[r] = aE_y[r] = aE_c[r] = aK_x[r] = aK_y[r] = aK_z[r] = aP_x[r] = aP_y[r] = aP_z[r] = 0.0;
    aC_x[r] = aC_y[r] = aC_z[r] = aW_z[r] = aD_c[r] = aA_x[r] = aA_y[r] = aA_z[r] = 0.0;
  }

  const unsigned FULLMASK = 0xffffffffu;
  const bool lane_out = lane < 31 && i <= P.Nx && (SH == 0 || i >= 1);
  // per-row output weight (1 for cells this thread reports, 0 for overlap cells) and output offset
  double wrow[RY];
  long long orow[RY];
  // integral-only multiplicities: how many of this block's output cells a face / edge owned by (lane, row) touches
  double mx[RY], my[RY], mxy[RY];          // x faces & fcf edges, y faces & cff edges, ffc edges
  double wE[RY];                           // EPV: weight of the corner point (i+1, j+1)
#pragma unroll
  for (int r = 0; r < RY; ++r) {
    const int jr = jr0 + r, j = j0 + jr;
    const bool row_out = jr < G::BR - 1 && j <= P.jy_hi && (SH == 0 || j >= P.jy_lo);
    const bool south_out = jr >= 1 && (jr - 1) < G::BR - 1 && (j - 1) <= P.jy_hi && (SH == 0 || (j - 1) >= P.jy_lo);   // the cell to the south, same lane
    wE[r] = (H_EPV && lane < 31 && i + 1 <= P.Nx && jr < G::BR - 1 && j + 1 <= P.jy_hi) ? 1.0 : 0.0;
    wrow[r] = (lane_out && row_out) ? 1.0 : 0.0;
    orow[r] = OUT ? (i * P.os[0] + j * P.os[1]) : 0;
    const double wsouth = (lane_out && south_out) ? 1.0 : 0.0;
    const double wwest = __shfl_up_sync(FULLMASK, wrow[r], 1);                              // the cell to the west, same row
    const double wsw = __shfl_up_sync(FULLMASK, wsouth, 1);
    mx[r] = wrow[r] + (lane > 0 ? wwest : 0.0);
    my[r] = wrow[r] + wsouth;
    mxy[r] = mx[r] + wsouth + (lane > 0 ? wsw : 0.0);
  }
  // OUT: one predicated accumulate (an FMA with the 0/1 weight) and one predicated store.
  // (Forming the store predicate once per row and plane as a boolean and the address outside the predicated region makes every
  //  store a bare predicated STG -- 18 % fewer instructions in the steady loop, no FP64 compare, no branch -- and measured SLOWER:
  //  29.4 against 25.5 ms at 1024^3 on the same board, SM clock up, i.e. the pipes idler; profiles/r2_summary.md.)
  auto emit = [&](int d, double raw, double scale, double w, int plane, int r) {
    acc[d] = fma(raw, w, acc[d]);
    if (OUT) {
      if (w != 0.0 && P.o[d].out) reinterpret_cast<TIN*>(P.o[d].out)[orow[r] + (long long)plane * P.os[2]] = (TIN)(raw * scale);
    }
  };

  int s = 0;
  uint32_t phase = 0;
  int par = 0;
  // One plane step.  STEADY (a compile-time tag): planes n, n-1 and n-2 all lie inside this block's z chunk, so the
  // plane-validity factors are the constants 1 and fold away; the few steps at the chunk ends take the general form.
  auto plane_step = [&](auto steady_tag, const int n) {
    constexpr bool STEADY = decltype(steady_tag)::value;
    mbar_wait(&full[s], phase);
    const TIN* T = tiles + s * (NF * TILE);
    const TIN* Tu = T, *Tv = T + TILE, *Tw = T + 2 * TILE, *Tb = T + F_B * TILE, *Tp = T + F_P * TILE;
    if (++s == NS) { s = 0; phase ^= 1; }
    // plane validity (uniform): cells of plane n, n-1, n-2 inside this block's z chunk
    const double pv0 = STEADY ? 1.0 : ((n >= k0 && n <= k1) ? 1.0 : 0.0), pv1 = STEADY ? 1.0 : (((n - 1) >= k0 && (n - 1) <= k1) ? 1.0 : 0.0),
                 pv2 = STEADY ? 1.0 : (((n - 2) >= k0 && (n - 2) <= k1) ? 1.0 : 0.0);
    // integral-only accumulation factors: terms of plane n, of plane n-1, of z face n (half-weights at chunk ends),
    // of z face n-1 -- all exactly 1 in the steady planes
    const double f0 = pv0, f1 = pv1, fz = 0.5 * (pv0 + pv1), fw = 0.5 * (pv1 + pv2);
    // mean-flow profiles at this plane and the two below (uniform loads; zero without profiles)
    double Un = 0.0, Un1 = 0.0, Un2 = 0.0, Vn = 0.0, Vn1 = 0.0, Vn2 = 0.0, Wn = 0.0, Wn1 = 0.0;
    if (PZ) {
      if (P.Uz) { Un = __ldg(P.Uz + n); Un1 = __ldg(P.Uz + n - 1); Un2 = __ldg(P.Uz + n - 2); }
      if (P.Vz) { Vn = __ldg(P.Vz + n); Vn1 = __ldg(P.Vz + n - 1); Vn2 = __ldg(P.Vz + n - 2); }
      if (P.Wz) { Wn = __ldg(P.Wz + n); Wn1 = __ldg(P.Wz + n - 1); }
    }
    // z metrics of this step (STR; otherwise the constants): face n, face n-1, cells n, n-1, n-2
    double rzf = rdz, rzf1 = rdz, rzc1 = rdz, dzc0 = 1.0, dzc1 = 1.0, dzc2 = 1.0, dzf0 = 1.0;
    if (STR) {
      const int lo = 1 - P.hz, hi = P.Nz + P.hz;
      const int q0 = min(max(n, lo), hi), q1 = min(max(n - 1, lo), hi), q2 = min(max(n - 2, lo), hi);
      rzf = __ldg(P.rdzf + q0); rzf1 = __ldg(P.rdzf + q1); rzc1 = __ldg(P.rdzc + q1);
      dzc0 = __ldg(P.dzc + q0); dzc1 = __ldg(P.dzc + q1); dzc2 = __ldg(P.dzc + q2); dzf0 = __ldg(P.dzf + q0);
    }
    const double ez_n = STR ? rzf * hsn : ez, dz2_1 = STR ? 2.0 * P.nu * (rzc1 * rzc1) : dz2;
    const bool ep = PZ && P.eps_pert;                   // dissipation of u - U: only the z differences change
    const double dUz = ep ? Un - Un1 : 0.0, dVz = ep ? Vn - Vn1 : 0.0, dWz = ep ? Wn - Wn1 : 0.0;
    auto accum = [&](double& a, double x, double y, double f) {     // a += f * x * y  (f folds away when STEADY)
      if (STEADY) a = fma(x, y, a); else a = fma(x * f, y, a);
    };

    // ---- stage 1: quantities owned by this thread's cells at plane n --------------------------------------
    double uc[RY], vc[RY], wc[RY], bc[RY], pc[RY];
    double X12[RY], P12[RY], P12E[RY], e23[RY], P23[RY], Y13[RY], P13[RY], P13E[RY], P13w[RY], P23w[RY];
    double Uu[RY], UuW[RY], Vv[RY], VvS[RY], inpl[RY], Kxy[RY], PRxy[RY], CHxy[RY], CDxy[RY], ucn[RY], vcn[RY];
    double XS23[RY], O13[RY], O12[RY], Dx[RY], Dy[RY], Sb[RY];
#pragma unroll
    for (int r = 0; r < RY; ++r) {
      const int oc = ro[r + 1], osr = ro[r], onr = ro[r + 2];   // this cell's tile row, the row to its south, to its north
      const TIN* ru = Tu + oc, *ruS = Tu + osr;                 // ru[0] west, ru[1] centre, ru[2] east
      const TIN* rv = Tv + oc, *rvS = Tv + osr, *rvN = Tv + onr;
      const TIN* rw = Tw + oc, *rwS = Tw + osr;
      uc[r] = ru[1]; vc[r] = rv[1]; wc[r] = rw[1];
      const double ue = ru[2], us = (r == 0) ? ruS[1] : uc[r > 0 ? r - 1 : 0];
      const double vw = rv[0], vn = rvN[1];
      const double ww = rw[0], ws = (r == 0) ? rwS[1] : wc[r > 0 ? r - 1 : 0];
      // ffc edge (i, j): scaled strain (nu/4 folded in) and the (Vu = Uv) momentum-flux product
      const double s12 = (uc[r] - us) * ey + (vc[r] - vw) * ex;
      P12[r] = (vw + vc[r]) * (us + uc[r]);
      // fcf edge (i, j, n) and cff edge
      const double s13 = (PZ ? (uc[r] - up[r]) - dUz : (uc[r] - up[r])) * ez_n + (wc[r] - ww) * ex;
      P13[r] = (ww + wc[r]) * (up[r] + uc[r]);
      const double s23 = (PZ ? (vc[r] - vp[r]) - dVz : (vc[r] - vp[r])) * ez_n + (wc[r] - ws) * ey;
      P23[r] = (ws + wc[r]) * (vp[r] + vc[r]);
      // w equation: the advecting transports are means of Ax u = dy dzc u (Ay v) over the two cell planes of the face
      P13w[r] = STR ? (ww + wc[r]) * fma(dzc1, up[r], dzc0 * uc[r]) : P13[r];
      P23w[r] = STR ? (ws + wc[r]) * fma(dzc1, vp[r], dzc0 * vc[r]) : P23[r];
      if (H_EPV) {
        // edge vorticities owned by this cell (always of the full velocities) and the b sums of the 2 x 2 columns
        // (i..i+1, j..j+1) of this plane
        O12[r] = (vc[r] - vw) * rdx - (uc[r] - us) * rdy;
        O13[r] = (uc[r] - up[r]) * rzf - (wc[r] - ww) * rdx;               // (rzf: 1 / dz_f of face n; rdz on a regular axis)
        const double o23 = (wc[r] - ws) * rdy - (vc[r] - vp[r]) * rzf;
        XS23[r] = o23 + __shfl_down_sync(FULLMASK, o23, 1);
        const TIN* rb = Tb + oc, *rbN = Tb + onr;
        const double b00 = rb[1], b10 = rb[2], b01 = rbN[1], b11 = rbN[2];
        Dx[r] = (b10 - b00) + (b11 - b01);
        Dy[r] = (b01 - b00) + (b11 - b10);
        Sb[r] = (b00 + b10) + (b01 + b11);
      }
      if (OUT) {
        const double e12 = s12 * s12, e13 = s13 * s13;
        e23[r] = s23 * s23;
        X12[r] = e12 + __shfl_down_sync(FULLMASK, e12, 1);      // x neighbours' edges by shuffle
        Y13[r] = e13 + __shfl_down_sync(FULLMASK, e13, 1);
      } else if (H_EPS) {
        // each edge term straight into its class sum (weights mxy, 2 mx, 2 my applied after the loop)
        accum(aE_xy[r], s12, s12, f0);
        accum(aE_x[r], s13, s13, fz);
        accum(aE_y[r], s23, s23, fz);
      }
      if (H_ADV) {
        P12E[r] = __shfl_down_sync(FULLMASK, P12[r], 1);
        P13E[r] = __shfl_down_sync(FULLMASK, P13w[r], 1);
        Uu[r] = (uc[r] + ue) * (uc[r] + ue);
        const double uw = ru[0];
        UuW[r] = (uw + uc[r]) * (uw + uc[r]);
        Vv[r] = (vc[r] + vn) * (vc[r] + vn);
        if (r == 0) { const double vs = rvS[1]; VvS[r] = (vs + vc[r]) * (vs + vc[r]); }
        else VvS[r] = Vv[r > 0 ? r - 1 : 0];
      }
      if (OUT) {
        if (H_SP) { ucn[r] = 0.5 * (uc[r] + ue) - Un; vcn[r] = 0.5 * (vc[r] + vn) - Vn; }   // raw tiles are recycled after the barrier
        if (H_EPS) {
          const double ax = ue - uc[r], ay = vn - vc[r];
          inpl[r] = fma(ax * ax, dx2, (ay * ay) * dy2);
        }
        if (H_K) Kxy[r] = fma(uc[r], uc[r], ue * ue) + fma(vc[r], vc[r], vn * vn);
        if (NEED_P) {
          const TIN* rp = Tp + oc;
          pc[r] = rp[1];
          const double pw = rp[0], pe = rp[2], ps = Tp[osr + 1], pn = Tp[onr + 1];
          const double tx_ = fma(ue, pe - pc[r], uc[r] * (pc[r] - pw));
          const double ty_ = fma(vn, pn - pc[r], vc[r] * (pc[r] - ps));
          PRxy[r] = fma(ty_, rdy, tx_ * rdx);
        }
        if (NEED_B) {
          const TIN* rb = Tb + oc;
          bc[r] = rb[1];
          const double dw_ = bc[r] - rb[0], de = rb[2] - bc[r], ds = bc[r] - Tb[osr + 1], dn = Tb[onr + 1] - bc[r];
          if (H_CHI) CHxy[r] = fma(fma(dw_, dw_, de * de), rdx2, fma(ds, ds, dn * dn) * rdy2);
          if (H_CDF) CDxy[r] = fma(dw_ - de, rdx2, (ds - dn) * rdy2);
        }
      } else {
        // integrals only: every face / cell term of plane n and of z face n goes into its class sum here
        const double dwz = PZ ? (wc[r] - wp[r]) - dWz : (wc[r] - wp[r]);
        if (H_SP) {
          // shear production of the cell below (plane n-1) about the profiles: -(u'w' dU/dz + v'w' dV/dz + w'w' dW/dz)
          const double wq1 = wc[r] - Wn, wq0 = wp[r] - Wn1;
          const double wcen = 0.5 * (wq0 + wq1), ww2 = 0.5 * fma(wq0, wq0, wq1 * wq1);
          const double spv = -(fma(ucp[r] * wcen, 0.5 * (Un - Un2) * rdz, fma(vcp[r] * wcen, 0.5 * (Vn - Vn2) * rdz, ww2 * ((Wn - Wn1) * rdz))));
          accum(aS_c[r], spv, 1.0, f1);
          ucp[r] = 0.5 * (uc[r] + ue) - Un;
          vcp[r] = 0.5 * (vc[r] + vn) - Vn;
        }
        if (H_EPS) {
          const double ax = ue - uc[r], ay = vn - vc[r];
          accum(aE_c[r], ax * dx2, ax, f0);
          accum(aE_c[r], ay * dy2, ay, f0);
          accum(aE_c[r], dwz * dz2, dwz, f1);               // S33 of the cell below
        }
        if (H_K) {
          accum(aK_x[r], uc[r], uc[r], f0);
          accum(aK_y[r], vc[r], vc[r], f0);
          accum(aK_z[r], wc[r], wc[r], fz);
        }
        if (NEED_P) {
          const TIN* rp = Tp + oc;
          pc[r] = rp[1];
          accum(aP_x[r], uc[r], pc[r] - rp[0], f0);
          accum(aP_y[r], vc[r], pc[r] - Tp[osr + 1], f0);
          accum(aP_z[r], wc[r], pc[r] - pp[r], fz);
        }
        if (NEED_B) {
          const TIN* rb = Tb + oc;
          bc[r] = rb[1];
          const double dw_ = bc[r] - rb[0], ds = bc[r] - Tb[osr + 1], dbz = bc[r] - bp[r];
          if (H_WB) accum(aW_z[r], wc[r], bp[r] + bc[r], fz);
          if (H_CHI) {
            accum(aC_x[r], dw_, dw_, f0);
            accum(aC_y[r], ds, ds, f0);
            accum(aC_z[r], dbz, dbz, fz);
          }
          if (H_CDF) {
            const double de = rb[2] - bc[r], dn = Tb[onr + 1] - bc[r];
            const double cdxy = fma(dw_ - de, rdx2, (ds - dn) * rdy2);
            const double lap = fma(-dbz, rdz2, CDacc[r]);   // Laplacian of the cell below, completed by this face
            accum(aD_c[r], bp[r], lap, f1);
            CDacc[r] = fma(dbz, rdz2, cdxy);
          }
        }
      }
    }

    // ---- y exchange between warps: row 0 of the warp to the north ---------------------------------------------
    double* myx = xch + ((par * NW + warp) * NX) * 32 + lane;
    if (H_ADV) { myx[0 * 32] = P12[0]; myx[1 * 32] = P23w[0]; }
    if (OUT) { myx[2 * 32] = X12[0]; myx[3 * 32] = e23[0]; myx[4 * 32] = Fvp[0]; }
    if (H_EPV) { myx[5 * 32] = XS23[0]; myx[6 * 32] = O13[0]; myx[7 * 32] = O12[0]; }
    __syncthreads();
    // every thread is done with raw stage s: refill it (plane n + NS)
    if (tid == 0 && n + NS <= n_last) issue(n + NS);   // same ring slot as plane n
    double X12N = 0.0, P12N = 0.0, e23N = 0.0, P23N = 0.0, FvN = 0.0, XS23N = 0.0, O13N = 0.0, O12N = 0.0;
    if (warp + 1 < NW) {
      const double* nx_ = xch + ((par * NW + warp + 1) * NX) * 32 + lane;
      if (H_ADV) { P12N = nx_[0 * 32]; P23N = nx_[1 * 32]; }
      if (OUT) { X12N = nx_[2 * 32]; e23N = nx_[3 * 32]; FvN = nx_[4 * 32]; }
      if (H_EPV) { XS23N = nx_[5 * 32]; O13N = nx_[6 * 32]; O12N = nx_[7 * 32]; }
    }
    // EPV is reported at z face n itself (no lag): faces k0..k1 of this chunk, plus the top face for the last chunk
    const double pvE = (STEADY || !H_EPV) ? 1.0 : ((n >= k0 && n <= k1 + ((k1 == P.kz_hi) ? P.epv_top : 0)) ? 1.0 : 0.0);

    // ---- stage 2 / finalisation, row by row (north neighbour: own row r+1, or the exchanged row) -----------------
#pragma unroll
    for (int r = 0; r < RY; ++r) {
      const bool last = (r == RY - 1);
      const double p12n = last ? P12N : P12[last ? r : r + 1];
      const double p23n = last ? P23N : P23w[last ? r : r + 1];
      const double dwz = PZ ? (wc[r] - wp[r]) - dWz : (wc[r] - wp[r]);
      if (H_EPV) {
        // q(i+1, j+1, n) = (f + ω)·∇b, every factor a 2- or 8-point mean onto the fff corner
        const double xs23n = last ? XS23N : XS23[last ? r : r + 1];
        const double o13n = last ? O13N : O13[last ? r : r + 1];
        const double o12n = last ? O12N : O12[last ? r : r + 1];
        const double ys13e = __shfl_down_sync(FULLMASK, O13[r] + o13n, 1);
        const double ne12 = __shfl_down_sync(FULLMASK, o12n, 1);
        const double qx = fma(0.5, xs23n, P.f[0]) * ((Dx[r] + Dxp[r]) * (0.25 * rdx));
        const double qy = fma(0.5, ys13e, P.f[1]) * ((Dy[r] + Dyp[r]) * (0.25 * rdy));
        const double qz = fma(0.5, ne12 + NE12p[r], P.f[2]) * ((Sb[r] - Sbp[r]) * (0.25 * rzf));
        const double raw = (qx + qy) + qz;
        Dxp[r] = Dx[r]; Dyp[r] = Dy[r]; Sbp[r] = Sb[r]; NE12p[r] = ne12;
        const double w = STR ? (pvE * wE[r]) * dzf0 : pvE * wE[r];          // (STR: the fff point's volume is dx dy dz_f)
        acc[FD_EPV] = fma(raw, w, acc[FD_EPV]);
        if (OUT) {
          if (w != 0.0 && P.o[FD_EPV].out) reinterpret_cast<TIN*>(P.o[FD_EPV].out)[orow[r] + P.os[0] + P.os[1] + (long long)n * P.os[2]] = (TIN)raw;
        }
      }
      if (OUT) {
        const double x12n = last ? X12N : X12[last ? r : r + 1];
        const double e23n = last ? e23N : e23[last ? r : r + 1];
        const double fvn = last ? FvN : Fvp[last ? r : r + 1];          // Fv(j+1) at plane n-2
        const double w1 = STR ? (pv1 * wrow[r]) * dzc1 : pv1 * wrow[r];   // plane n-1 terms (STR: the cell's dz_c is its volume weight)
        const double w2 = STR ? (pv2 * wrow[r]) * dzc2 : pv2 * wrow[r];   // ADV lags one more plane
        // eps(n-1) = [in-plane + xy edges](n-1) + z strain + the fcf/cff edge sums at faces n-1 and n
        if (H_EPS) {
          const double edge_n = Y13[r] + (e23[r] + e23n);
          double raw;
          if (!STR) {
            raw = fma(dwz * dwz, dz2, Eacc[r] + edge_n);
            Eacc[r] = (inpl[r] + edge_n) + (X12[r] + x12n);
          } else {
            const double ew = dzf0 * edge_n;                 // Az dz_f: the weight of an fcf / cff edge in its two cells
            raw = fma(dwz * dwz, dz2_1, fma(rzc1, Efp[r] + ew, Eacc[r]));
            Eacc[r] = inpl[r] + (X12[r] + x12n);
            Efp[r] = ew;
          }
          emit(FD_EPS, raw, SC_EPS, w1, n - 1, r);
        }
        if (H_K) {
          const double wsq = wc[r] * wc[r];
          const double raw = Kacc[r] + wsq;
          Kacc[r] = Kxy[r] + wsq;
          emit(FD_K, raw, SC_K, w1, n - 1, r);
        }
        if (H_PR) {
          const double fz = wc[r] * (pc[r] - pp[r]);
          const double raw = fma(fz, rzf, PRacc[r]);
          PRacc[r] = fma(fz, rzf, PRxy[r]);
          emit(FD_PR, raw, SC_PR, w1, n - 1, r);
        }
        if (NEED_B) {
          const double dbz = bc[r] - bp[r];
          if (H_WB) {
            const double hz = wc[r] * (bp[r] + bc[r]);
            const double raw = hzp[r] + hz;
            hzp[r] = hz;
            emit(FD_WB, raw, SC_WB, w1, n - 1, r);
          }
          if (H_CHI) {
            const double gz = dbz * dbz;
            double raw;
            if (!STR) {
              raw = fma(gz, rdz2, CHacc[r]);
              CHacc[r] = fma(gz, rdz2, CHxy[r]);
            } else {
              const double gw = gz * rzf;                    // Az dz_f (db/dz)^2 / (dx dy)
              raw = fma(rzc1, Gzp[r] + gw, CHacc[r]);
              CHacc[r] = CHxy[r];
              Gzp[r] = gw;
            }
            emit(FD_CHI, raw, SC_CHI, w1, n - 1, r);
          }
          if (H_CDF) {
            double raw;
            if (!STR) {
              raw = bp[r] * fma(-dbz, rdz2, CDacc[r]);
              CDacc[r] = fma(dbz, rdz2, CDxy[r]);
            } else {
              const double gl = dbz * rzf;                   // db/dz at face n
              raw = bp[r] * fma(rzc1, Gcp[r] - gl, CDacc[r]);
              CDacc[r] = CDxy[r];
              Gcp[r] = gl;
            }
            emit(FD_CDF, raw, SC_CDF, w1, n - 1, r);
          }
        }
        if (H_SP) {
          const double wq1 = wc[r] - Wn, wq0 = wp[r] - Wn1;
          const double wcen = 0.5 * (wq0 + wq1), ww2 = 0.5 * fma(wq0, wq0, wq1 * wq1);
          // mean shear of cell n-1 (STR: the mean of its two face gradients; dW/dz over the cell's dz_c)
          const double gU = STR ? 0.5 * fma(Un - Un1, rzf, (Un1 - Un2) * rzf1) : 0.5 * (Un - Un2) * rdz;
          const double gV = STR ? 0.5 * fma(Vn - Vn1, rzf, (Vn1 - Vn2) * rzf1) : 0.5 * (Vn - Vn2) * rdz;
          const double gW = (Wn - Wn1) * (STR ? rzc1 : rdz);
          const double raw = -(fma(ucp[r] * wcen, gU, fma(vcp[r] * wcen, gV, ww2 * gW)));
          ucp[r] = ucn[r];
          vcp[r] = vcn[r];
          emit(FD_SP, raw, 1.0, w1, n - 1, r);
        }
        if (H_ADV) {
          // faces of plane n-1: F = velocity x flux divergence (the common 1/8 is the diagnostic's scale factor)
          const double Fu = up[r] * fma(P13[r] - P13p[r], rzc1, dU[r]);
          const double Fv = vp[r] * fma(P23[r] - P23p[r], rzc1, dV[r]);
          const double wsum = wp[r] + wc[r];
          const double Wwp = wsum * wsum;                                 // Ww at cell plane n-1
          const double Fw = wp[r] * fma(Wwp - Wwpp[r], rzf1, dW[r]);     // face n-1
          // ADV(n-2) = 1/8 [Fu + FuE + Fv + FvN + Fw(n-2) + Fw(n-1)]
          const double raw = (Apart[r] + fvn) + Fw;
          const double FuE = __shfl_down_sync(FULLMASK, Fu, 1);
          Apart[r] = (Fu + FuE) + (Fv + Fw);
          Fvp[r] = Fv;
          Wwpp[r] = Wwp;
          dU[r] = fma(Uu[r] - UuW[r], rdx, (p12n - P12[r]) * rdy);
          dV[r] = fma(P12E[r] - P12[r], rdx, (Vv[r] - VvS[r]) * rdy);
          dW[r] = fma(P13E[r] - P13w[r], rdx, (p23n - P23w[r]) * rdy);
          if (STR) dW[r] *= rzf;                                          // / dz_f of face n (V_ccf)
          P13p[r] = P13[r]; P23p[r] = P23[r];
          emit(FD_ADV, raw, SC_ADV, w2, n - 2, r);
        }
      } else {
        // ---- integrals only: what is left after the exchange is the advection faces of plane n-1 ---------------------
        if (H_ADV) {
          const double wsum = wp[r] + wc[r];
          const double Wwp = wsum * wsum;
          accum(aA_x[r], up[r], fma(P13[r] - P13p[r], rdz, dU[r]), f1);
          accum(aA_y[r], vp[r], fma(P23[r] - P23p[r], rdz, dV[r]), f1);
          accum(aA_z[r], wp[r], fma(Wwp - Wwpp[r], rdz, dW[r]), fw);      // z face n-1: cells n-1 and n-2
          Wwpp[r] = Wwp;
          dU[r] = fma(Uu[r] - UuW[r], rdx, (p12n - P12[r]) * rdy);
          dV[r] = fma(P12E[r] - P12[r], rdx, (Vv[r] - VvS[r]) * rdy);
          dW[r] = fma(P13E[r] - P13w[r], rdx, (p23n - P23w[r]) * rdy);
          P13p[r] = P13[r]; P23p[r] = P23[r];
        }
      }
      up[r] = uc[r]; vp[r] = vc[r]; wp[r] = wc[r];
      if (NEED_B) bp[r] = bc[r];
      if (NEED_P) pp[r] = pc[r];
    }
    par ^= 1;
  };
  {
    int n = n_first;
    const int steady_lo = k0 + 2, steady_hi = k1;
    for (; n <= n_last && n < steady_lo; ++n) plane_step(std::false_type{}, n);
    if (OUT) {
      // two planes per trip: the pipeline registers of the first become the "previous plane" of the second by renaming
      // instead of by 64-bit moves
#pragma unroll 2
      for (; n <= steady_hi; ++n) plane_step(std::true_type{}, n);
    } else {
      for (; n <= steady_hi; ++n) plane_step(std::true_type{}, n);
    }
    for (; n <= n_last; ++n) plane_step(std::false_type{}, n);
  }

  if (!OUT) {
    // apply the multiplicity of each class once: x faces / fcf edges mx, y faces / cff edges my, ffc edges mxy,
    // cells wrow, z faces (shared by the cells above and below) 2 wrow
#pragma unroll
    for (int r = 0; r < RY; ++r) {
      const double w2 = 2.0 * wrow[r];
      if (H_EPS) acc[FD_EPS] += mxy[r] * aE_xy[r] + 2.0 * (mx[r] * aE_x[r] + my[r] * aE_y[r]) + wrow[r] * aE_c[r];
      if (H_K) acc[FD_K] += mx[r] * aK_x[r] + my[r] * aK_y[r] + w2 * aK_z[r];
      if (H_PR) acc[FD_PR] += rdx * (mx[r] * aP_x[r]) + rdy * (my[r] * aP_y[r]) + rdz * (w2 * aP_z[r]);
      if (H_WB) acc[FD_WB] += w2 * aW_z[r];
      if (H_CHI) acc[FD_CHI] += rdx2 * (mx[r] * aC_x[r]) + rdy2 * (my[r] * aC_y[r]) + rdz2 * (w2 * aC_z[r]);
      if (H_CDF) acc[FD_CDF] += wrow[r] * aD_c[r];
      if (H_ADV) acc[FD_ADV] += mx[r] * aA_x[r] + my[r] * aA_y[r] + w2 * aA_z[r];
      if (H_SP) acc[FD_SP] += wrow[r] * aS_c[r];
    }
  }
  // ---- volume integrals: warp shuffle, then one partial per block and slot -------------------------------------
  if (P.nslots > 0) {
#pragma unroll
    for (int d = 0; d < FD_N; ++d) {
      if (!(MASK & (1 << d)) || P.o[d].slot < 0) continue;
      double v = acc[d];
      for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(FULLMASK, v, o);
      __syncthreads();
      if (lane == 0) red[warp] = v;
      __syncthreads();
      if (tid == 0) {
        double s = 0.0;
        for (int w = 0; w < NW; ++w) s += red[w];
        const double scale[FD_N] = {SC_K, SC_ADV, SC_PR, SC_WB, SC_EPS, SC_CHI, SC_CDF, 1.0, 1.0, 1.0, 1.0};
        P.partial[(size_t)blockIdx.x * P.nslots + P.o[d].slot] = s * (P.V * scale[d]);
      }
    }
  }
}

#include "ocb_budget_sums.cuh"

}  // namespace

struct ocb_fused_plan {
  FusedParams P;
  CUtensorMap maps[6];             // u, v, w, b, p, nu_e
  int mask, nw, ry, ns;
  bool sums;                       // ocb_budget_sums_kernel (integral-only class sums) instead of ocb_budget_kernel
  int minb;                        // resident blocks per SM the sums kernel was compiled for
  int tw;                          // tile row pitch in elements (the TMA box width)
  int nblocks;
  size_t smem;
  void (*kernel)(FusedParams, CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap);
  bool nuf;                        // the eddy-viscosity variant of the class-sum kernel (nu_e as sixth input)
  bool f32;                        // Float32-input plan
  double* ztab_d;                  // Float32 plans on a stretched axis: Float64 copies of the z-metric tables
  // Float32 plans about mean profiles: the profiles' Float32 arrays (whole parent extents) and their Float64 copies, refreshed at every execute
  const float* prof_src[3];
  double* prof_d;
  int prof_n, prof_len[3];
};

template <int MASK, int NW, int RY, int NS>
static int setup_variant(ocb_fused_plan* fp, bool out) {
  using G = Geo<NW, RY>;
  constexpr bool NEED_B = MASK & ((1 << FD_WB) | (1 << FD_CHI) | (1 << FD_CDF) | (1 << FD_EPV)), NEED_P = MASK & (1 << FD_PR);
  constexpr int NF = 3 + (NEED_B ? 1 : 0) + (NEED_P ? 1 : 0);
  fp->nw = NW; fp->ry = RY; fp->ns = NS; fp->mask = MASK;
  fp->smem = sizeof(double) * ((size_t)NS * NF * G::TILE + 2 * NW * NXCH * 32 + NW) + sizeof(uint64_t) * NS;
  // ask for the largest shared-memory carveout so that occupancy is set by registers, not by the default split
  if (out) {
    cudaFuncSetAttribute(ocb_budget_kernel<MASK, NW, RY, NS, true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    fp->kernel = ocb_budget_kernel<MASK, NW, RY, NS, true>;
    OCB_CUDA(cudaFuncSetAttribute(ocb_budget_kernel<MASK, NW, RY, NS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fp->smem));
  } else {
    cudaFuncSetAttribute(ocb_budget_kernel<MASK, NW, RY, NS, false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    fp->kernel = ocb_budget_kernel<MASK, NW, RY, NS, false>;
    OCB_CUDA(cudaFuncSetAttribute(ocb_budget_kernel<MASK, NW, RY, NS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fp->smem));
  }
  return OCB_OK;
}

// stretched z: the all-terms field / cell-integral kernel with per-plane metrics (always the OUT = true form; a plan without
// output fields simply has no destinations)
template <int MASK, int NW, int RY, int NS>
static int setup_variant_stretched(ocb_fused_plan* fp) {
  const int rc = setup_variant<MASK, NW, RY, NS>(fp, true);       // shape bookkeeping and shared-memory size are the same
  if (rc != OCB_OK) return rc;
  auto kern = ocb_budget_kernel<MASK, NW, RY, NS, true, true>;
  cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  fp->kernel = kern;
  OCB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fp->smem));
  return OCB_OK;
}

// Float32 fields in and out: the cell-value form with two-box tiles (40-column pitch)
template <int MASK, int NW, int RY, int NS, bool STR>
static int setup_variant_f32(ocb_fused_plan* fp) {
  using G = Geo<NW, RY>;
  constexpr bool NEED_B = MASK & ((1 << FD_WB) | (1 << FD_CHI) | (1 << FD_CDF)), NEED_P = MASK & (1 << FD_PR);
  constexpr int NF = 3 + (NEED_B ? 1 : 0) + (NEED_P ? 1 : 0);
  constexpr int TILE_F = 2 * (((40 * (G::TR / 2) * 4 + 127) / 128) * 128 / 4);
  fp->nw = NW; fp->ry = RY; fp->ns = NS; fp->mask = MASK; fp->tw = 40;
  fp->smem = sizeof(float) * (size_t)NS * NF * TILE_F + sizeof(double) * (2 * NW * NXCH * 32 + NW) + sizeof(uint64_t) * NS;
  auto kern = ocb_budget_kernel<MASK, NW, RY, NS, true, STR, float>;
  cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  fp->kernel = kern;
  OCB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fp->smem));
  return OCB_OK;
}

template <int MASK, int NW, int RY, int NS, int MINB, int TWK = TW, bool STR = false, bool NUF = false, class TIN = double>
static int setup_sums(ocb_fused_plan* fp) {
  using G = Geo<NW, RY>;
  constexpr bool NEED_B = MASK & ((1 << FD_WB) | (1 << FD_CHI) | (1 << FD_CDF)), NEED_P = MASK & (1 << FD_PR);
  constexpr int NF = 3 + (NEED_B ? 1 : 0) + (NEED_P ? 1 : 0) + (NUF ? 1 : 0);
  fp->nw = NW; fp->ry = RY; fp->ns = NS; fp->mask = MASK; fp->sums = true; fp->minb = MINB; fp->tw = TWK; fp->nuf = NUF;
  constexpr int TILE_E = sizeof(TIN) == 4 ? 2 * sums_tile_elems_of(TWK, G::TR / 2, 4) : sums_tile_elems_of(TWK, G::TR, (int)sizeof(TIN));
  fp->smem = sizeof(TIN) * (size_t)NS * NF * TILE_E + sizeof(double) * NW + sizeof(uint64_t) * NS;
  auto kern = ocb_budget_sums_kernel<MASK, NW, RY, NS, MINB, TWK, STR, NUF, TIN>;
  cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  fp->kernel = kern;
  OCB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fp->smem));
  return OCB_OK;
}
// the default shapes in both tile pitches: 34 columns when every tile starts on a 16-byte boundary (odd x halo)
template <int MASK, int NW, int RY, int NS, int MINB, bool STR = false, bool NUF = false>
static int setup_sums_pitch(ocb_fused_plan* fp, bool narrow) {
  return narrow ? setup_sums<MASK, NW, RY, NS, MINB, 34, STR, NUF>(fp) : setup_sums<MASK, NW, RY, NS, MINB, 36, STR, NUF>(fp);
}

static int make_map(EncodeTiledFn enc, const ocb_field& f, int rows, CUtensorMap* map, int tw = TW) {
  cuuint64_t dims[3] = {(cuuint64_t)f.size[0], (cuuint64_t)f.size[1], (cuuint64_t)f.size[2]};
  cuuint64_t strides[2] = {(cuuint64_t)f.stride[1] * 8, (cuuint64_t)f.stride[2] * 8};
  cuuint32_t box[3] = {(cuuint32_t)tw, (cuuint32_t)rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, f.data, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) OCB_FAIL(OCB_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return OCB_OK;
}

static int fused_diag_index(int diag) {
  switch (diag) {
    case OCB_DIAG_KE: return FD_K;
    case OCB_DIAG_ADV: return FD_ADV;
    case OCB_DIAG_PR: return FD_PR;
    case OCB_DIAG_WB: return FD_WB;
    case OCB_DIAG_EPS: return FD_EPS;
    case OCB_DIAG_CHI: return FD_CHI;
    case OCB_DIAG_CDIFF: return FD_CDF;
    case OCB_DIAG_SP: case OCB_DIAG_SPZ: return FD_SP;   // equal for z-profile means (the x and y parts vanish)
    case OCB_DIAG_EPV: return FD_EPV;
    case OCB_DIAG_TKE: return FD_TKE;                    // class-sum kernel only (integrals about mean profiles)
    default: return -1;
  }
}
bool ocb_fused_handles_diag(int diag) { return fused_diag_index(diag) >= 0; }

// a mean-flow operand usable by the profile kernels: ZeroField (-> nullptr) or a field reduced over x and y whose z
// axis is contiguous; *out is pre-offset so that out[k] is the value at the 1-based level k
static bool profile_pointer(const ocb_field& f, int zloc, const double** out) {
  *out = nullptr;
  if (f.kind == OCB_FIELD_ZERO || (f.kind == OCB_FIELD_ARRAY && !f.data)) return true;
  if (f.kind != OCB_FIELD_ARRAY) return false;
  if (f.loc[0] != OCB_NOTHING || f.loc[1] != OCB_NOTHING || f.loc[2] != zloc || f.stride[2] != 1) return false;
  *out = (const double*)f.data + (f.halo[2] - 1);
  return true;
}

__global__ void ocb_f32_to_f64_kernel(const float* __restrict__ a, double* __restrict__ b, int n) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) b[t] = (double)a[t];
}

static bool tma_compatible(const ocb_field& f, const ocb_grid* g) {
  if (f.kind != OCB_FIELD_ARRAY || !f.data) return false;
  for (int d = 0; d < 3; ++d) if (f.loc[d] == OCB_NOTHING || f.halo[d] != g->H[d]) return false;
  if (f.stride[0] != 1) return false;
  if (((uintptr_t)f.data & 15) || ((f.stride[1] * 8) & 15) || ((f.stride[2] * 8) & 15)) return false;
  if (f.stride[1] != f.size[0] || f.stride[2] != (long long)f.size[0] * f.size[1]) return false;   // dense parent
  return true;
}

// T = float: only integral-only plans of the seven budget terms, on the Float32-input form of the class-sum kernel
template <class T>
static int fused_try_create_t(const ocb_grid* grid, const ocb_sweep_inputs* in, const ocb_request* reqs, int n,
                              const SweepParams<T>& SP, ocb_fused_plan** out) {
  constexpr bool F32 = sizeof(T) == 4;
  auto tma_compatible = [](const ocb_field& f, const ocb_grid* g) { return F32 ? tma_compatible_f32(f, g) : ::tma_compatible(f, g); };
  *out = nullptr;
  if (getenv("OCB_DISABLE_FUSED")) return OCB_OK;
  if (F32 && getenv("OCB_DISABLE_SUMS_F32")) return OCB_OK;
  // ---- envelope ------------------------------------------------------------------------------------------
  // stretched z: the seven budget terms (no mean-flow operands, no potential vorticity) on the per-plane-metric variant
  const bool stretched = grid->dzc || grid->dzf;
  if (stretched && (!SP.g.dzc || !SP.g.dzf || !SP.g.rdzc || !SP.g.rdzf || getenv("OCB_DISABLE_FUSED_STRETCHED"))) return OCB_OK;
  if (grid->H[0] < 2 || grid->H[1] < 2 || grid->H[2] < 2) return OCB_OK;
  // eddy-viscosity / diffusivity fields: only the closure-independent terms (K, ADV, PR, w b) -- checked below, once the mask is known
  const bool closure_fields = in->closure.n_nu_fields || in->closure.n_kappa_fields;
  int mask = 0;
  int which_out[FD_N], which_slot[FD_N];     // a diagnostic may appear twice: once as a Field, once as an Integral
  for (int d = 0; d < FD_N; ++d) which_out[d] = which_slot[d] = -1;
  bool eps_pert = false, any_pert = false, any_outq = false;
  int sp_diag = -1, cell_req = -1, epv_top = 0;
  // The class-sum kernel (integrals only, x periodic, no potential vorticity) takes the dissipation of u and of u - U side by
  // side (FD_EPS / FD_EPSP) and the turbulent kinetic energy; the cell-exact kernel one dissipation request and no TKE.
  for (int r = 0; r < n; ++r) any_outq = any_outq || SP.req[r].out;
  // (stretched z: the class-sum form is compiled for all terms -- five inputs -- and for the TKE budget with pressure
  //  redistribution -- u, v, w, p; plans whose operands do not cover the variant they would need stay cell-exact)
  bool str_sums_ok = false;
  if (stretched && !getenv("OCB_DISABLE_SUMS_STRETCHED")) {
    bool only_tkep = true;
    for (int r = 0; r < n; ++r) {
      const int dg = reqs[r].diag;
      const bool pert = reqs[r].flags & OCB_REQ_PERTURBATION;
      if (!(dg == OCB_DIAG_SP || dg == OCB_DIAG_SPZ || dg == OCB_DIAG_TKE || dg == OCB_DIAG_PR || dg == OCB_DIAG_KE || (dg == OCB_DIAG_EPS && pert)))
        only_tkep = false;
    }
    str_sums_ok = tma_compatible(in->p, grid) && (only_tkep || tma_compatible(in->b, grid));
  }
  bool sums_convention = !any_outq && grid->topo[0] == OCB_PERIODIC && !getenv("OCB_FUSED_CELL_INTEGRALS") && (!stretched || str_sums_ok);
  for (int r = 0; r < n; ++r) if (reqs[r].diag == OCB_DIAG_EPV) sums_convention = false;
  if (F32 && !sums_convention && getenv("OCB_DISABLE_FIELDS_F32")) return OCB_OK;
  for (int r = 0; r < n; ++r) {
    int d = fused_diag_index(reqs[r].diag);
    if (d < 0) return OCB_OK;
    const bool pert = reqs[r].flags & OCB_REQ_PERTURBATION;
    if (d == FD_TKE) {
      if (!sums_convention) return OCB_OK;              // (always about U, V, W: src/TurbulentKineticEnergyEquation.jl:24-30)
      any_pert = true;
    } else if (d == FD_EPV) {
      // the corner-owned fff term: periodic x and y, H >= 3, and the whole z extent (Nz + 1 faces on a Bounded axis)
      if (pert || grid->topo[0] != OCB_PERIODIC || grid->topo[1] != OCB_PERIODIC) return OCB_OK;
      if (grid->H[0] < 3 || grid->H[1] < 3 || grid->H[2] < 3) return OCB_OK;
      epv_top = grid->topo[2] == OCB_BOUNDED ? 1 : 0;
      if (SP.req[r].lo[2] != 1 || SP.req[r].hi[2] != grid->N[2] + epv_top) return OCB_OK;
    } else if (d == FD_SP) {
      if (!pert) return OCB_OK;                       // u' handed over as fields: generic kernel
      if (sp_diag >= 0 && sp_diag != reqs[r].diag) return OCB_OK;
      sp_diag = reqs[r].diag;
    } else if (d == FD_EPS) {
      if (sums_convention) {
        if (pert) d = FD_EPSP;
      } else {
        if (which_out[d] >= 0 || which_slot[d] >= 0) { if (pert != eps_pert) return OCB_OK; }
        eps_pert = pert;
      }
    } else if (pert) {
      return OCB_OK;
    }
    any_pert = any_pert || pert;
    const DReq& R = SP.req[r];
    if (R.lo[0] != 1 || R.hi[0] != grid->N[0]) return OCB_OK;                                               // full x extent
    if (R.lo[1] != SP.req[0].lo[1] || R.hi[1] != SP.req[0].hi[1]) return OCB_OK;                            // one y window
    if (d != FD_EPV) {                                                                                        // one z window
      if (cell_req < 0) cell_req = r;
      else if (R.lo[2] != SP.req[cell_req].lo[2] || R.hi[2] != SP.req[cell_req].hi[2]) return OCB_OK;
    }
    if (R.out) { if (which_out[d] >= 0) return OCB_OK; which_out[d] = r; }
    if (R.slot >= 0) { if (which_slot[d] >= 0) return OCB_OK; which_slot[d] = r; }
    mask |= 1 << d;
  }
  // Float32: the seven plain budget terms, or the TKE budget with pressure redistribution about mean profiles
  constexpr int TKEP_F = (1 << FD_SP) | (1 << FD_EPSP) | (1 << FD_TKE) | (1 << FD_PR) | (1 << FD_K);
  if (F32 && (any_pert ? (mask & ~TKEP_F) != 0 : (mask & ~FD_BUDGET) != 0)) return OCB_OK;
  if (F32 && !sums_convention && (any_pert || (mask & ~FD_BUDGET))) return OCB_OK;      // with output fields: the seven plain terms
  // stretched z: the cell-exact variant evaluates the seven plain budget terms; the class-sum form also the TKE-budget terms
  constexpr int TKE_TERMS = (1 << FD_SP) | (1 << FD_EPSP) | (1 << FD_TKE);
  // (cell-exact: the seven plain terms, or the TKE budget about mean profiles {eps(u - U), SP, PR, K} -- no potential vorticity)
  constexpr int TKEM_S = (1 << FD_EPS) | (1 << FD_SP) | (1 << FD_PR) | (1 << FD_K);
  if (stretched && !sums_convention) {
    const int rest = mask & ~(1 << FD_EPV);
    if (mask & (1 << FD_EPV)) { if (rest & ~TKEM_S) return OCB_OK; }              // the potential vorticity alone or with the TKE budget
    else if (any_pert ? (rest & ~TKEM_S) != 0 : (rest & ~FD_BUDGET) != 0) return OCB_OK;
  }
  if (stretched && sums_convention && (mask & ~(FD_BUDGET | TKE_TERMS))) return OCB_OK;
  // ... and, as class sums, the dissipation with ONE eddy-viscosity field (the nu_e of SmagorinskyLilly / AMD as a sixth input)
  constexpr int CLOSURE_TERMS = (1 << FD_EPS) | (1 << FD_EPSP) | (1 << FD_CHI) | (1 << FD_CDF);
  bool nuf = false;
  if (closure_fields && (mask & CLOSURE_TERMS)) {
    const ocb_field& nf = in->closure.nu_fields[0];
    const bool nu_ok = sums_convention && in->closure.n_nu_fields == 1 && tma_compatible(nf, grid) &&
                       nf.loc[0] == OCB_CENTER && nf.loc[1] == OCB_CENTER && nf.loc[2] == OCB_CENTER && !getenv("OCB_DISABLE_SUMS_NUF");
    // the tracer-variance terms: a constant kappa, or kappa_e = scale x the SAME field (SmagorinskyLilly's nu_e / Pr)
    const ocb_field& kf = in->closure.kappa_fields[0];
    const bool kappa_ok = in->closure.n_kappa_fields == 0 ||
                          (in->closure.n_kappa_fields == 1 && kf.data == nf.data && kf.stride[1] == nf.stride[1] && kf.stride[2] == nf.stride[2]);
    const int kappa_terms = (1 << FD_CHI) | (1 << FD_CDF);
    if (!nu_ok || ((mask & kappa_terms) && !kappa_ok)) return OCB_OK;   // velocity-gradient / generic kernels
    nuf = true;
  }
  if ((mask & (1 << FD_EPV)) && cell_req >= 0 && (SP.req[cell_req].lo[2] != 1 || SP.req[cell_req].hi[2] != grid->N[2])) return OCB_OK;
  const double *Uz = nullptr, *Vz = nullptr, *Wz = nullptr;
  if (any_pert) {
    // perturbations about horizontal-mean profiles only (Average(u, dims=(1,2)) reduced fields, or ZeroField)
    if (grid->H[2] < 3) return OCB_OK;
    if (!profile_pointer(in->U, OCB_CENTER, &Uz) || !profile_pointer(in->V, OCB_CENTER, &Vz) || !profile_pointer(in->W, OCB_FACE, &Wz)) return OCB_OK;
    if (!sums_convention) mask |= 1 << FD_SP;         // cell-exact kernel: the profile-carrying variants are the ones with the SP term
  }
  if ((mask & (1 << FD_WB)) && (in->ghat[0] != 0.0 || in->ghat[1] != 0.0)) return OCB_OK;   // vertical gravity only
  bool any_out = false;
  long long os[3] = {0, 0, 0};
  for (int d = 0; d < FD_N; ++d) {
    if (which_out[d] < 0) continue;
    const DReq& R = SP.req[which_out[d]];
    if (any_out && (R.os[0] != os[0] || R.os[1] != os[1] || R.os[2] != os[2])) return OCB_OK;   // one layout for all outputs
    any_out = true;
    for (int a = 0; a < 3; ++a) os[a] = R.os[a];
  }
  const bool need_b = mask & ((1 << FD_WB) | (1 << FD_CHI) | (1 << FD_CDF) | (1 << FD_EPV)), need_p = mask & (1 << FD_PR);
  if (!tma_compatible(in->u, grid) || !tma_compatible(in->v, grid) || !tma_compatible(in->w, grid)) return OCB_OK;
  if (need_b && !tma_compatible(in->b, grid)) return OCB_OK;
  if (need_p && !tma_compatible(in->p, grid)) return OCB_OK;
  if (in->u.loc[0] != OCB_FACE || in->v.loc[1] != OCB_FACE || in->w.loc[2] != OCB_FACE) return OCB_OK;
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return OCB_OK;

  ocb_fused_plan* fp = new (std::nothrow) ocb_fused_plan();
  if (!fp) OCB_FAIL(OCB_ERR_ALLOC, "out of host memory");
  memset((void*)fp, 0, sizeof(*fp));
  // ---- variant: all seven terms, or the dissipation-only kernel ---------------------------------------------
  constexpr int ALL = FD_BUDGET;
  int rc;
  // Shapes measured on B200 at 1024^3 (profiles/r1_summary.md).  One 256-thread block per SM with the full 255
  // registers and no spills beats every higher-occupancy shape: integrals only -> 8 warps x 3 rows (99 Gcells/s vs 90
  // for 12 x 2), with output fields -> 8 warps x 2 rows (44 vs 29 Gcells/s for 6 x 2).
  // OCB_FUSED_CONFIG="NW,RY,NS" overrides the shape of the all-terms kernel for tuning runs.
  const char* cfg = getenv("OCB_FUSED_CONFIG");
  int nw = 8, ry = any_out ? 2 : 3, ns = 3;
  if (cfg) sscanf(cfg, "%d,%d,%d", &nw, &ry, &ns);
  constexpr int VEL = (1 << FD_K) | (1 << FD_ADV) | (1 << FD_EPS);                 // velocity-only terms
  constexpr int VELP = VEL | (1 << FD_PR);                                          // + pressure
  constexpr int VELB = VEL | (1 << FD_WB) | (1 << FD_CHI) | (1 << FD_CDF);          // + tracer
  constexpr int KAPW = (1 << FD_K) | (1 << FD_ADV) | (1 << FD_PR) | (1 << FD_WB);   // K, ADV, PR, w b
#define OCB_PICK(M) (any_out ? setup_variant<M, 8, 2, 3>(fp, true) : setup_variant<M, 8, 3, 3>(fp, false))
  constexpr int TKEM = (1 << FD_EPS) | (1 << FD_SP) | (1 << FD_PR) | (1 << FD_K);  // TKE budget about mean profiles
  constexpr int EPVM = 1 << FD_EPV, TKEV = TKEM | EPVM;                            // + Ertel PV (BASELINE config 2)
  // Integrals only, x periodic, no perturbation operands: the class-sum kernel (ocb_budget_sums.cuh).
  // OCB_FUSED_CELL_INTEGRALS=1 keeps the cell-exact integral mode of ocb_budget_kernel (the two are compared in the tests);
  // OCB_SUMS_CONFIG="NW,RY,NS,MINB" picks another compiled shape of the all-terms kernel for tuning runs.
  const bool sums = sums_convention;
  const int njy = SP.req[0].hi[1] - SP.req[0].lo[1] + 1;
  if constexpr (F32) {
    // Float32 inputs: the all-terms class sums in the two-box tile form (40-column pitch), with the stretched / eddy-viscosity forms;
    // plans with output fields (or a Bounded x): the cell-value kernel, all seven terms
    if (!sums) {
      if (!tma_compatible(in->b, grid) || !tma_compatible(in->p, grid)) { delete fp; return OCB_OK; }
      rc = stretched ? setup_variant_f32<ALL, 8, 2, 3, true>(fp) : setup_variant_f32<ALL, 8, 2, 3, false>(fp);
    } else if (any_pert) {
      if (!tma_compatible(in->p, grid)) { delete fp; return OCB_OK; }
      if (stretched) rc = nuf ? setup_sums<TKEP_F, 4, 4, 3, 2, 40, true, true, float>(fp) : setup_sums<TKEP_F, 4, 4, 3, 2, 40, true, false, float>(fp);
      else rc = nuf ? setup_sums<TKEP_F, 4, 4, 3, 2, 40, false, true, float>(fp) : setup_sums<TKEP_F, 4, 4, 3, 2, 40, false, false, float>(fp);
    } else {
      if (!tma_compatible(in->b, grid) || !tma_compatible(in->p, grid)) { delete fp; return OCB_OK; }
      if (stretched) rc = nuf ? setup_sums<ALL, 4, 4, 3, 2, 40, true, true, float>(fp) : setup_sums<ALL, 4, 4, 3, 2, 40, true, false, float>(fp);
      else rc = nuf ? setup_sums<ALL, 4, 4, 3, 2, 40, false, true, float>(fp) : setup_sums<ALL, 4, 4, 3, 2, 40, false, false, float>(fp);
    }
  } else {
  if (sums && nuf) {
    // eddy-viscosity dissipation: the all-terms class sums with nu_e inside (needs all five other inputs as well)
    const bool narrow = (grid->H[0] & 1) && !getenv("OCB_SUMS_WIDE_TILES");
    constexpr int TKEB_N = (1 << FD_SP) | (1 << FD_EPSP) | (1 << FD_TKE), TKEP_N = TKEB_N | (1 << FD_PR) | (1 << FD_K);
    if (mask & TKEB_N) {
      // about mean profiles: the TKE budget with pressure redistribution {SP, eps(u - U), PR, K / TKE} (u, v, w, p, nu_e)
      if ((mask & ~TKEP_N) || !tma_compatible(in->p, grid)) { delete fp; return OCB_OK; }
      rc = stretched ? setup_sums_pitch<TKEP_N, 4, 4, 3, 2, true, true>(fp, narrow) : setup_sums_pitch<TKEP_N, 4, 4, 3, 2, false, true>(fp, narrow);
    } else {
      if (!tma_compatible(in->b, grid) || !tma_compatible(in->p, grid)) { delete fp; return OCB_OK; }
      rc = stretched ? setup_sums_pitch<ALL, 4, 4, 3, 2, true, true>(fp, narrow) : setup_sums_pitch<ALL, 4, 4, 3, 2, false, true>(fp, narrow);
    }
  } else if (stretched && sums) {
    const bool narrow = (grid->H[0] & 1) && !getenv("OCB_SUMS_WIDE_TILES");
    constexpr int TKEB_S = (1 << FD_SP) | (1 << FD_EPSP) | (1 << FD_TKE);
    constexpr int TKEP_S = TKEB_S | (1 << FD_PR) | (1 << FD_K);
    if (mask & TKEB_S) {
      if ((mask & ~TKEP_S) == 0) rc = setup_sums_pitch<TKEP_S, 4, 4, 3, 2, true>(fp, narrow);
      else rc = setup_sums_pitch<ALL | TKEB_S, 4, 4, 3, 2, true>(fp, narrow);
    }
    else if (njy + 1 <= 4) rc = setup_sums_pitch<ALL, 2, 2, 3, 4, true>(fp, narrow);
    else rc = setup_sums_pitch<ALL, 4, 4, 3, 2, true>(fp, narrow);
  } else if (stretched) {
    if (mask & EPVM) rc = mask == EPVM ? setup_variant_stretched<EPVM, 8, 2, 3>(fp) : setup_variant_stretched<TKEV, 8, 2, 3>(fp);
    else if (mask & (1 << FD_SP)) rc = setup_variant_stretched<TKEM, 8, 2, 3>(fp);
    else if ((mask & ~VEL) == 0) rc = setup_variant_stretched<VEL, 8, 2, 3>(fp);
    else if ((mask & ~VELP) == 0) rc = setup_variant_stretched<VELP, 8, 2, 3>(fp);
    else if ((mask & ~VELB) == 0) rc = setup_variant_stretched<VELB, 8, 2, 3>(fp);
    else rc = setup_variant_stretched<ALL, 8, 2, 3>(fp);
  } else if (sums) {
    const char* sc = getenv("OCB_SUMS_CONFIG");
    int a = 0, b = 0, c = 0, d = 0;
    if (sc) sscanf(sc, "%d,%d,%d,%d", &a, &b, &c, &d);
    const int key = a * 1000 + b * 100 + c * 10 + d;
    const bool narrow = (grid->H[0] & 1) && !getenv("OCB_SUMS_WIDE_TILES");   // every tile's first column is 16-byte aligned
    constexpr int TKEB = (1 << FD_SP) | (1 << FD_EPSP) | (1 << FD_TKE);              // the TKE-budget terms about mean profiles
    constexpr int ALLT = ALL | TKEB;                                                  // TKE + KE + tracer-variance budget: 10 integrals
    constexpr int TKEP = TKEB | (1 << FD_PR) | (1 << FD_K);                           // TKE budget with pressure redistribution (config 2 without EPV)
    if (mask & TKEB) {
      if ((mask & ~TKEP) == 0) rc = setup_sums_pitch<TKEP, 4, 4, 3, 2>(fp, narrow);
      else rc = setup_sums_pitch<ALLT, 4, 4, 3, 2>(fp, narrow);
    }
    else if (!sc && njy + 1 <= 4) rc = setup_sums_pitch<ALL, 2, 2, 3, 4>(fp, narrow);          // slab-edge strips: 4-row tiles
    else if (!sc && mask == (1 << FD_EPS)) rc = setup_sums_pitch<(1 << FD_EPS), 8, 3, 3, 1>(fp, narrow);
    else if (!sc && (mask & ~VEL) == 0) rc = setup_sums_pitch<VEL, 8, 3, 3, 1>(fp, narrow);
    else if (!sc && (mask & ~VELP) == 0) rc = setup_sums_pitch<VELP, 8, 3, 3, 1>(fp, narrow);
    else if (!sc && (mask & ~VELB) == 0) rc = setup_sums_pitch<VELB, 8, 3, 3, 1>(fp, narrow);
    // the closure-independent terms (what the budget kernels keep of a plan with eddy-viscosity / diffusivity fields)
    else if (!sc && (mask & ~KAPW) == 0) rc = setup_sums_pitch<KAPW, 4, 4, 3, 2>(fp, narrow);
    else if (key == 8331) rc = setup_sums<ALL, 8, 3, 3, 1>(fp);
    else if (key == 8341) rc = setup_sums<ALL, 8, 3, 4, 1>(fp);
    else if (key == 8431) rc = setup_sums<ALL, 8, 4, 3, 1>(fp);
    else if (key == 12231) rc = setup_sums<ALL, 12, 2, 3, 1>(fp);
    else if (key == 4432) rc = setup_sums<ALL, 4, 4, 3, 2>(fp);
    else if (key == 4442) rc = setup_sums<ALL, 4, 4, 4, 2>(fp);
    else if (key == 4532) rc = setup_sums<ALL, 4, 5, 3, 2>(fp);
    else if (key == 4631) rc = setup_sums_pitch<ALL, 4, 6, 3, 1>(fp, narrow);
    else rc = setup_sums_pitch<ALL, 4, 4, 3, 2>(fp, narrow);               // measured best at 1024^3 (profiles/r1_summary.md)
  } else if (mask & EPVM) {
    if (mask == EPVM) rc = OCB_PICK(EPVM);
    else if ((mask & ~TKEV) == 0) rc = OCB_PICK(TKEV);
    else { delete fp; return OCB_OK; }
  } else if (mask & (1 << FD_SP)) {
    if (mask & ~TKEM) { delete fp; return OCB_OK; }
    rc = OCB_PICK(TKEM);
  } else if (mask == (1 << FD_EPS)) rc = OCB_PICK((1 << FD_EPS));
  else if ((mask & ~VEL) == 0) rc = OCB_PICK(VEL);
  else if ((mask & ~VELP) == 0) rc = OCB_PICK(VELP);
  else if ((mask & ~VELB) == 0) rc = OCB_PICK(VELB);
  else if (!cfg && !any_out && SP.req[0].hi[1] - SP.req[0].lo[1] + 1 <= 3)
    // a strip of up to three rows (the slab-edge sweeps of the overlapped multi-GPU step): 4-row tiles instead of 24-row ones
    rc = setup_variant<ALL, 2, 2, 3>(fp, false);
  else if (!cfg) rc = OCB_PICK(ALL);
#undef OCB_PICK
  else if (nw == 16 && ry == 1) rc = setup_variant<ALL, 16, 1, 3>(fp, any_out);
  else if (nw == 8 && ry == 3) rc = setup_variant<ALL, 8, 3, 3>(fp, any_out);
  else if (nw == 6 && ry == 2) rc = setup_variant<ALL, 6, 2, 3>(fp, any_out);
  else if (nw == 8 && ry == 2) rc = setup_variant<ALL, 8, 2, 3>(fp, any_out);
  else rc = setup_variant<ALL, 12, 2, 3>(fp, any_out);
  }
  if (rc != OCB_OK) { delete fp; return rc; }
  const bool vneed_b = fp->mask & ((1 << FD_WB) | (1 << FD_CHI) | (1 << FD_CDF) | (1 << FD_EPV)), vneed_p = fp->mask & (1 << FD_PR);
  if ((vneed_b && !tma_compatible(in->b, grid)) || (vneed_p && !tma_compatible(in->p, grid))) { delete fp; return OCB_OK; }

  FusedParams& P = fp->P;
  P.Nx = grid->N[0]; P.Ny = grid->N[1]; P.Nz = grid->N[2];
  P.hx = grid->H[0]; P.hy = grid->H[1]; P.hz = grid->H[2];
  const int BR = fp->nw * fp->ry;
  const int sh = (fp->mask & (1 << FD_EPV)) ? 1 : 0;          // corner-owned fff points: the tiling starts one cell early
  P.tiles_x = (P.Nx + sh + 30) / 31;
  P.jy_lo = SP.req[0].lo[1]; P.jy_hi = SP.req[0].hi[1];
  P.tiles_y = ((P.jy_hi - P.jy_lo + 1) + sh + BR - 2) / (BR - 1);
  if (fp->sums) {                                              // disjoint 32 x BR tiles over rows jy_lo .. jy_hi + 1
    // (.. jy_hi when the window is the whole periodic y extent, or a slab whose halo rows hold the ring neighbours' rows:
    //  the items of row jy_hi + 1 are then row jy_lo's of the wrapped domain / next slab and are counted there)
    P.y_ring = (grid->topo[1] == OCB_PERIODIC && P.jy_lo == 1 && P.jy_hi == grid->N[1] && !getenv("OCB_SUMS_NO_RING")) ? 1 : 0;
    P.tiles_x = (P.Nx + 31) / 32;
    P.tiles_y = (njy + (P.y_ring ? 0 : 1) + BR - 1) / BR;
  }
  const int tiles = P.tiles_x * P.tiles_y;
  P.kz_lo = cell_req >= 0 ? SP.req[cell_req].lo[2] : 1;
  P.kz_hi = cell_req >= 0 ? SP.req[cell_req].hi[2] : grid->N[2];
  P.epv_top = epv_top;
  for (int a = 0; a < 3; ++a) P.f[a] = in->f[a];
  const int nz_win = P.kz_hi - P.kz_lo + 1;
  int nchunks = (ocb_sm_count() * 8 + tiles - 1) / tiles;
  if (nchunks < 1) nchunks = 1;
  int kchunk = (nz_win + nchunks - 1) / nchunks;
  if (kchunk < 32) kchunk = nz_win < 32 ? nz_win : 32;
  if (fp->sums) {
    // z chunks of the class-sum kernel: the chunk count with the fewest block waves x planes per block (each chunk
    // re-reads two warm-up planes; the last wave of a launch is partly empty)
    const long long slots = (long long)ocb_sm_count() * fp->minb;
    long long best = -1;
    for (int nc = 1; nc <= 16; ++nc) {
      int kc = (nz_win + nc - 1) / nc;
      if (kc < 32) kc = nz_win < 32 ? nz_win : 32;
      const int ncr = (nz_win + kc - 1) / kc;
      const long long waves = ((long long)tiles * ncr + slots - 1) / slots;
      const long long cost = waves * (kc + 3) * 16 + nc;          // ties: fewer chunks
      if (best < 0 || cost < best) { best = cost; kchunk = kc; }
    }
  }
  if (const char* kc = getenv("OCB_FUSED_KCHUNK")) kchunk = atoi(kc) > 0 ? atoi(kc) : kchunk;
  P.kchunk = kchunk;
  P.nchunks = (nz_win + kchunk - 1) / kchunk;
  fp->nblocks = tiles * P.nchunks;
  P.rdx = 1.0 / grid->d[0]; P.rdy = 1.0 / grid->d[1]; P.rdz = stretched ? 1.0 : 1.0 / grid->d[2];
  P.nu = in->closure.nu_const; P.kappa = in->closure.kappa_const; P.ghat_z = in->ghat[2];
  P.kappa_scale = in->closure.n_kappa_fields ? in->closure.kappa_scale[0] : 0.0;
  P.V = stretched ? grid->d[0] * grid->d[1] : grid->d[0] * grid->d[1] * grid->d[2];      // stretched: dz_c is the cell's integral weight
  P.rowlen = in->u.size[0];
  fp->f32 = F32;
  if constexpr (F32) {
    P.dzc = P.dzf = P.rdzc = P.rdzf = nullptr;
    if (stretched) {
      // Float64 copies of the four metric tables (contiguous in the plan's Float32 table block, each nz entries, pre-offset by Hz - 1)
      const int nz = grid->N[2] + 2 * grid->H[2] + 1, off = grid->H[2] - 1;
      OCB_CUDA(cudaMalloc(&fp->ztab_d, sizeof(double) * 4 * nz));
      ocb_f32_to_f64_kernel<<<(4 * nz + 127) / 128, 128>>>(SP.g.dzc - off, fp->ztab_d, 4 * nz);
      OCB_CUDA(cudaGetLastError());
      P.dzc = fp->ztab_d + 0 * nz + off; P.dzf = fp->ztab_d + 1 * nz + off; P.rdzc = fp->ztab_d + 2 * nz + off; P.rdzf = fp->ztab_d + 3 * nz + off;
    }
  } else {
    P.dzc = SP.g.dzc; P.dzf = SP.g.dzf; P.rdzc = SP.g.rdzc; P.rdzf = SP.g.rdzf;
  }
  P.nslots = SP.nslots;
  P.Uz = Uz; P.Vz = Vz; P.Wz = Wz; P.eps_pert = eps_pert ? 1 : 0;
  if constexpr (F32) {
    if (any_pert) {
      // the profiles are Float32 arrays that change from step to step: Float64 copies, refreshed by every execute
      const int np_ = grid->N[2] + 2 * grid->H[2] + 1;
      OCB_CUDA(cudaMalloc(&fp->prof_d, sizeof(double) * 3 * np_));
      OCB_CUDA(cudaMemset(fp->prof_d, 0, sizeof(double) * 3 * np_));
      fp->prof_n = np_;
      const ocb_field* pf[3] = {&in->U, &in->V, &in->W};
      const double** dst[3] = {&P.Uz, &P.Vz, &P.Wz};
      for (int a = 0; a < 3; ++a) {
        fp->prof_src[a] = nullptr;
        if (*dst[a]) {                                   // (profile_pointer accepted it: an array reduced over x and y, contiguous in z)
          fp->prof_src[a] = (const float*)pf[a]->data;
          if (pf[a]->size[2] > np_) { delete fp; return OCB_OK; }
          fp->prof_len[a] = pf[a]->size[2];
          *dst[a] = fp->prof_d + a * np_ + (pf[a]->halo[2] - 1);
        }
      }
    }
  }
  for (int a = 0; a < 3; ++a) P.os[a] = os[a];
  for (int d = 0; d < FD_N; ++d) {
    P.o[d].out = nullptr; P.o[d].slot = -1;
    if (which_out[d] >= 0) P.o[d].out = (double*)SP.req[which_out[d]].out;
    if (which_slot[d] >= 0) P.o[d].slot = SP.req[which_slot[d]].slot;
  }
  const ocb_field* f[6] = {&in->u, &in->v, &in->w, vneed_b ? &in->b : &in->u, vneed_p ? &in->p : &in->u, fp->nuf ? &in->closure.nu_fields[0] : &in->u};
  for (int a = 0; a < 6; ++a) {
    rc = F32 ? make_map_f32(enc, *f[a], BR + 2, &fp->maps[a], fp->tw) : make_map(enc, *f[a], BR + 2, &fp->maps[a], fp->sums ? fp->tw : TW);
    if (rc != OCB_OK) { delete fp; return rc; }
  }
  *out = fp;
  return OCB_OK;
}

int ocb_fused_try_create(const ocb_grid* grid, const ocb_sweep_inputs* in, const ocb_request* reqs, int n,
                         const SweepParams<float>& SP, ocb_fused_plan** out) {
  return fused_try_create_t<float>(grid, in, reqs, n, SP, out);
}
int ocb_fused_try_create(const ocb_grid* grid, const ocb_sweep_inputs* in, const ocb_request* reqs, int n,
                         const SweepParams<double>& SP, ocb_fused_plan** out) {
  return fused_try_create_t<double>(grid, in, reqs, n, SP, out);
}

int ocb_fused_nblocks(const ocb_fused_plan* fp) { return fp->nblocks; }

// (the eddy-viscosity variant reads a sixth input whose halo rows the ring step does not push, and the ring step is only
//  exercised on Float64 slabs: neither is ring-capable)
bool ocb_fused_ring_capable(const ocb_fused_plan* fp) { return fp->sums && fp->P.y_ring && !fp->nuf && !fp->f32; }

int ocb_fused_execute_gated(ocb_fused_plan* fp, double* partial, long long* flags, long long step, long long timeout_cycles, cudaStream_t st) {
  OCB_REQUIRE(ocb_fused_ring_capable(fp), "gated execution needs the class-sum kernel over a whole periodic y extent");
  fp->P.gate = flags; fp->P.gate_step = step; fp->P.gate_timeout = timeout_cycles;
  const int rc = ocb_fused_execute(fp, partial, st);
  fp->P.gate = nullptr;
  return rc;
}

int ocb_fused_execute(ocb_fused_plan* fp, double* partial, cudaStream_t st) {
  fp->P.partial = partial;
  if (fp->prof_d) {
    for (int a = 0; a < 3; ++a)
      if (fp->prof_src[a]) ocb_f32_to_f64_kernel<<<(fp->prof_len[a] + 127) / 128, 128, 0, st>>>(fp->prof_src[a], fp->prof_d + a * fp->prof_n, fp->prof_len[a]);
  }
  dim3 block(32, fp->nw, 1), grid(fp->nblocks, 1, 1);
  fp->kernel<<<grid, block, fp->smem, st>>>(fp->P, fp->maps[0], fp->maps[1], fp->maps[2], fp->maps[3], fp->maps[4], fp->maps[5]);
  OCB_CUDA(cudaGetLastError());
  return OCB_OK;
}

void ocb_fused_destroy(ocb_fused_plan* fp) {
  if (fp && fp->ztab_d) cudaFree(fp->ztab_d);
  if (fp && fp->prof_d) cudaFree(fp->prof_d);
  delete fp;
}
