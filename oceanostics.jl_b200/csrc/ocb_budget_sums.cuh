// K1a-sums: the integral-only form of the budget kernel (included by ocb_sweep_fused.cu, inside its anonymous namespace).
//
// When a plan asks only for volume integrals of the seven budget terms
//   K, ADV, PR, WB, EPS, CHI, CDF   (reference lines: see the table at the top of ocb_sweep_fused.cu)
// over (all of x) x (a y window) x (a z window) of an x-periodic grid, the sum over cells of every term is re-associated
// into sums over the staggered items the cell values are built from -- x/y/z faces, ffc corners, fcf / cff edges, cells --
// each item weighted by the number of WINDOW cells it belongs to (class weights that depend on the row j and the plane k
// only, never on the lane).  Every block owns a disjoint (32 x BR) set of columns: no overlap rows or lanes, no shuffles,
// no exchange between warps, and -- away from the window edges -- one set of accumulators per thread.
//
// The advection term is summed by parts: with the face-located momentum fluxes
//   Uu = (u_i + u_{i+1})^2 (ccc), P12 = (v_{i-1} + v_i)(u_{j-1} + u_j) (ffc), P13 = (w_{i-1} + w_i)(u_{k-1} + u_k) (fcf), ...
// the reference's  sum_cells 1/8 [Fu + FuE + Fv + FvN + Fw + FwT],  Fu = u (dx Uu/Dx + dy P12/Dy + dz P13/Dz)  etc.
// (src/KineticEnergyEquation.jl:147-152 with Oceananigans' Centered(order=2) div_Uu/div_Uv/div_Uw) equals EXACTLY (a
// finite re-association of the same products; x periodic, window-edge faces keep their explicit boundary terms)
//   -1/4 sum [ P12 S12 + P13 S13 + P23 S23 + Uu dxu + Vv dyv + Ww dzw ]
// over interior items, where S12 = dyu + dxv, S13 = dzu + dxw, S23 = dzv + dyw are the strain sums the dissipation rate
// needs anyway.  Rows and planes at a window edge take the general weighted form (GEN below); everything else the FAST
// form.  Per cell: 68 Float64 instructions for the seven integrals (the cell-exact integral mode of ocb_budget_kernel: ~100).
//
// Ownership in z: a block sweeps the planes k0..k1 of its chunk (plus the two closing steps of the window for the last
// chunk); the step at plane n handles the cell, face, corner and edge items of plane n and the z-centred items of the cell
// below (S33, Ww, the tracer Laplacian).  Chunks other than the first pre-load two planes to warm that one-plane lag.

// TWK = tile row pitch in elements: 34 (west halo + 32 + east halo) when the tile's first column is 16-byte aligned for every
// tile (odd x halo), else 36 (one spare column either side of the aligned box start, as in ocb_budget_kernel).
constexpr int sums_tile_elems(int tw, int tr) { return ((tw * tr * 8 + 127) / 128) * 128 / 8; }   // 128-byte aligned field tiles
constexpr int sums_tile_elems_of(int tw, int tr, int esize) { return ((tw * tr * esize + 127) / 128) * 128 / esize; }

// STR = true: z is a stretched axis (the seven budget terms only).  The class sums then carry per-plane weights -- the cell's dz_c
// for the items of a cell plane, dz_f of the face for the fcf / cff strain edges, (dz_c(n-1) + dz_c(n)) / 2 for the z-face
// items of K and w b, 1 / dz_f and 1 / dz_c where a squared z difference stays divided once after the volume weight cancels one
// factor -- and the by-parts advection splits its fcf / cff products in two: the u (v) equation's z part, whose cell weight
// cancels the 1 / dz_c of the divergence, and the w equation's x (y) part with the dz_c-weighted advecting transport (the same
// forms as ocb_budget_kernel<.., STR>, re-associated; oracle/kernels.py has the reference's area / volume weighted formulas).
// NUF = true: the closure carries one eddy-viscosity field nu_e (ccc; SmagorinskyLilly / AMD) next to the constant part P.nu.  The
// dissipation's items then carry their own viscosity -- nu at ccc for the diagonal strains, the four-point means of nu_e at
// ffc / fcf / cff for the edges (Oceananigans' interpolation of the closure field, as K1c and the oracle form it) -- inside the
// class sums instead of one factor after the loop; nu_e is a sixth input tile.
// TIN = float: the INPUTS are Float32 fields (the sums, like every integral of the library, are Float64).  A Float32 parent row
// is not a 16-byte multiple in general (Nx + 2H = 774 floats), which TMA cannot stride over; two rows are (Ny + 2H even), so the
// tensor maps describe each plane as rows of 2 (Nx + 2H) elements and a tile comes as TWO boxes per field -- the parent rows of
// one parity and those of the other, each with its own 16-byte-aligned start column -- into the two halves of the tile.  Every
// access to the tile goes through the per-row offsets `ro[]`; nothing else of the kernel changes (TWK = 40 for this form).
template <int MASK, int NW, int RY, int NS, int MINB, int TWK, bool STR = false, bool NUF = false, class TIN = double>
__global__ void __launch_bounds__(NW * 32, MINB)
ocb_budget_sums_kernel(const __grid_constant__ FusedParams P, const __grid_constant__ CUtensorMap mu, const __grid_constant__ CUtensorMap mv,
                       const __grid_constant__ CUtensorMap mw, const __grid_constant__ CUtensorMap mb, const __grid_constant__ CUtensorMap mp,
                       const __grid_constant__ CUtensorMap mn) {
  using G = Geo<NW, RY>;
  constexpr bool F32IN = sizeof(TIN) == 4;
  // (Float32: each of the two boxes of a tile starts on its own 128-byte boundary, as TMA destinations must)
  constexpr int HALF = sums_tile_elems_of(TWK, G::TR / 2, (int)sizeof(TIN));
  constexpr int TILE = F32IN ? 2 * HALF : sums_tile_elems_of(TWK, G::TR, (int)sizeof(TIN));
  static_assert(!F32IN || G::TR % 2 == 0, "the two-box form needs an even number of tile rows");
  constexpr bool H_K = MASK & (1 << FD_K), H_ADV = MASK & (1 << FD_ADV), H_PR = MASK & (1 << FD_PR), H_WB = MASK & (1 << FD_WB),
                 H_EPS = MASK & (1 << FD_EPS), H_CHI = MASK & (1 << FD_CHI), H_CDF = MASK & (1 << FD_CDF),
                 H_SP = MASK & (1 << FD_SP), H_EPSP = MASK & (1 << FD_EPSP), H_TKE = MASK & (1 << FD_TKE);
  // PZ: terms about horizontal-mean profiles U(z), V(z), W(z) (the TKE budget): shear production (its z part -- the x and y
  // parts vanish for profiles), the dissipation rate of u - U (only the z differences of the strains change, by per-plane
  // constants) and the turbulent kinetic energy
  constexpr bool PZ = H_SP || H_EPSP || H_TKE;
  constexpr bool NEED_B = H_WB || H_CHI || H_CDF, NEED_P = H_PR, NEED_S = H_EPS || H_ADV || H_EPSP;
  constexpr int NF = 3 + (NEED_B ? 1 : 0) + (NEED_P ? 1 : 0) + (NUF ? 1 : 0);
  constexpr int F_B = 3, F_P = NEED_B ? 4 : 3, F_N = NF - 1;
  static_assert(!NUF || H_EPS || H_EPSP, "the eddy-viscosity variant needs a dissipation term");

  extern __shared__ __align__(1024) double smem_d[];
  TIN* tiles = reinterpret_cast<TIN*>(smem_d);                                 // [NS][NF][TILE]
  double* red = reinterpret_cast<double*>(tiles + NS * NF * TILE);             // [NW]
  uint64_t* full = reinterpret_cast<uint64_t*>(red + NW);                      // [NS]

  const int lane = threadIdx.x, warp = threadIdx.y;
  const int tid = warp * 32 + lane;
  int bid = blockIdx.x;
  const int tx = bid % P.tiles_x; bid /= P.tiles_x;
  int ty, chunk;
  if (P.gate) {
    // slab of a ring (ocb_plan_execute_ring): the two tile rows that read a neighbour's halo row come LAST in launch
    // order, behind every tile row that needs no neighbour data, and wait for that neighbour's push of this step below
    chunk = bid % P.nchunks;
    const int t = bid / P.nchunks;
    ty = P.tiles_y <= 2 ? t : (t < P.tiles_y - 2 ? t + 1 : (t == P.tiles_y - 2 ? 0 : P.tiles_y - 1));
  } else {
    ty = bid % P.tiles_y;
    chunk = bid / P.tiles_y;
  }
  const int i0 = 1 + 32 * tx, j0 = P.jy_lo + G::BR * ty;
  const int k0 = P.kz_lo + chunk * P.kchunk;
  const int k1 = min(k0 + P.kchunk - 1, P.kz_hi);
  const int n_first = chunk == 0 ? k0 - 1 : k0 - 2;
  const int n_last = (k1 == P.kz_hi) ? k1 + (H_ADV ? 2 : 1) : k1;
  const int i = i0 + lane;
  const int jr0 = warp * RY;

  // ---- TMA ring (tiles of TWK x (BR + 2) elements per field, 16-byte aligned box start) -------------------------
  constexpr uint32_t STAGE_BYTES = (uint32_t)(NF * TWK * G::TR * sizeof(TIN));
  const int px_raw = i0 - 1 + P.hx - 1, py = j0 - 1 + P.hy - 1;
  const int xoff = px_raw & 1, px = px_raw - xoff;
  // Float32 inputs: box b holds the tile rows b, b + 2, .. (parent rows py + b, py + b + 2, ..: one parity), read from the
  // merged-row view at column (parity) * rowlen + px_raw, rounded down to a 16-byte boundary (4 elements)
  int bx[2] = {0, 0}, by[2] = {0, 0}, bo[2] = {0, 0};
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
    if (!F32IN) {
      tma_load_3d(dst + 0 * TILE, &mu, &full[s], px, py, pz);
      tma_load_3d(dst + 1 * TILE, &mv, &full[s], px, py, pz);
      tma_load_3d(dst + 2 * TILE, &mw, &full[s], px, py, pz);
      if (NEED_B) tma_load_3d(dst + F_B * TILE, &mb, &full[s], px, py, pz);
      if (NEED_P) tma_load_3d(dst + F_P * TILE, &mp, &full[s], px, py, pz);
      if (NUF) tma_load_3d(dst + F_N * TILE, &mn, &full[s], px, py, pz);
    } else {
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        TIN* half = dst + b * HALF;
        tma_load_3d(half + 0 * TILE, &mu, &full[s], bx[b], by[b], pz);
        tma_load_3d(half + 1 * TILE, &mv, &full[s], bx[b], by[b], pz);
        tma_load_3d(half + 2 * TILE, &mw, &full[s], bx[b], by[b], pz);
        if (NEED_B) tma_load_3d(half + F_B * TILE, &mb, &full[s], bx[b], by[b], pz);
        if (NEED_P) tma_load_3d(half + F_P * TILE, &mp, &full[s], bx[b], by[b], pz);
        if (NUF) tma_load_3d(half + F_N * TILE, &mn, &full[s], bx[b], by[b], pz);
      }
    }
  };
  if (tid == 0) {
    for (int s = 0; s < NS; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  bool gate_timed_out = false;
  if (tid == 0) {
    if (P.gate && (ty == 0 || ty == P.tiles_y - 1)) {
      // flags[0]: the low neighbour's push has landed in my south halo row, flags[1]: the high neighbour's in my north one
      const long long t0 = clock64();
      const bool need_lo = ty == 0, need_hi = ty == P.tiles_y - 1;
      for (;;) {
        long long flo = P.gate_step, fhi = P.gate_step;
        if (need_lo) asm volatile("ld.acquire.sys.global.s64 %0, [%1];" : "=l"(flo) : "l"(P.gate) : "memory");
        if (need_hi) asm volatile("ld.acquire.sys.global.s64 %0, [%1];" : "=l"(fhi) : "l"(P.gate + 1) : "memory");
        if (flo >= P.gate_step && fhi >= P.gate_step) break;
        if (clock64() - t0 > P.gate_timeout) { gate_timed_out = true; P.gate[2] = 1; break; }   // reported: flag word + NaN partials
        __nanosleep(100);
      }
      asm volatile("fence.proxy.async;" ::: "memory");   // the rows were written through the generic proxy (peer stores); TMA reads them
    }
    for (int n = n_first; n < n_first + NS && n <= n_last; ++n) issue(n);
  }

  const double rdx = P.rdx, rdy = P.rdy, rdz = P.rdz;
  const double rdx2 = rdx * rdx, rdy2 = rdy * rdy, rdz2 = rdz * rdz;
  // tile offsets of this lane's cell in the thread's rows, the row to the south (ro[0]) and the row to the north (ro[RY + 1])
  int ro[RY + 2];
#pragma unroll
  for (int q = 0; q < RY + 2; ++q) {
    const int tr = jr0 + q;                                  // tile row (row 0 = the south halo row of the block)
    ro[q] = F32IN ? (tr & 1) * HALF + (tr >> 1) * TWK + bo[tr & 1] + lane + 1 : tr * TWK + lane + xoff + 1;
  }

  // ---- per-row classes (functions of j only, warp-uniform) ----------------------------------------------------------
  // alpha = [row j in the window], beta = [row j-1 in it], alphan = [row j+1 in it]; rows jy_lo .. jy_hi+1 carry items
  const bool xfull = i0 + 31 <= P.Nx;
  const bool lane_ok = i <= P.Nx;
  bool rlive[RY];
#pragma unroll
  // P.y_ring: the y window is a whole periodic extent, or one slab of a periodic ring of slabs.  Then every row's items
  // carry the interior weights (the items of row jy_hi + 1 ARE those of the next slab's / the wrapped row jy_lo, whose
  // boundary terms they cancel), so rows jy_lo .. jy_hi are all FAST rows and no block works on the extra row.
  const bool ring = P.y_ring != 0;
  for (int r = 0; r < RY; ++r) rlive[r] = j0 + jr0 + r <= P.jy_hi + (ring ? 0 : 1);
  // the whole warp is away from the window's y edges (rows jy_lo+1 .. jy_hi-1) in a full x tile: FAST steps
  const bool tfast = xfull && (ring ? (j0 + jr0 + RY - 1 <= P.jy_hi) : (j0 + jr0 >= P.jy_lo + 1 && j0 + jr0 + RY - 1 <= P.jy_hi - 1));

  // pipeline state: the centre values of this plane and of the previous one in two register sets that swap roles from
  // step to step (the plane loop is unrolled by two, so nothing is copied), and the in-plane + lower-face part of the
  // tracer Laplacian
  double cu[2][RY], cv[2][RY], cw[2][RY], cb[2][RY], cp[2][RY], CDacc[RY];
  double cn[2][RY], xn[2][RY], yn[2][RY];                 // NUF: nu_e at the cell, and its sums with the west / south neighbour
#pragma unroll
  for (int r = 0; r < RY; ++r) {
    cu[0][r] = cv[0][r] = cw[0][r] = cb[0][r] = cp[0][r] = cu[1][r] = cv[1][r] = cw[1][r] = cb[1][r] = cp[1][r] = CDacc[r] = 0.0;
    cn[0][r] = cn[1][r] = xn[0][r] = xn[1][r] = yn[0][r] = yn[1][r] = 0.0;
  }
  // class sums, normalised to the interior weights (the closing lines of the kernel give each one its factor)
  double aEe = 0.0, aEx = 0.0, aEy = 0.0, aEz = 0.0;       // S12^2 + S13^2 + S23^2 | (dx u)^2 | (dy v)^2 | (dz w)^2   (undivided differences)
  double aAe = 0.0, aAx = 0.0, aAy = 0.0, aAz = 0.0;       // P12 S12 + P13 S13 + P23 S23 | Uu dxu | Vv dyv | Ww dzw
  double aK = 0.0;                                         // u^2 + v^2 + w^2 at their faces
  double aPx = 0.0, aPy = 0.0, aPz = 0.0;                  // u dx p | v dy p | w dz p
  double aW = 0.0;                                         // w (b_{k-1} + b_k)
  double aCx = 0.0, aCy = 0.0, aCz = 0.0, aD = 0.0;        // (dx b)^2 | (dy b)^2 | (dz b)^2 | b lap(b)
  double aEeP = 0.0, aEzP = 0.0;                           // the same as aEe, aEz for the strains of u - U
  double aKP = 0.0;                                        // (u-U)^2 + (v-V)^2 + (w-W)^2 at their faces
  // shear production, summed by faces: sum_cells ubar' wbar' GU_k = 1/4 sum_{x faces, z faces n} (w'_{i-1} + w'_i)_n (GU_n u'_n + GU_{n-1} u'_{n-1})
  // (x periodic; ubar', wbar' the two-point means to the cell centre, GU_k = (U_{k+1} - U_{k-1})/2 the undivided mean shear at cell k),
  // likewise for v with the y faces, and sum_cells (w'^2)bar GW_k = 1/2 sum_{z faces n} w'_n^2 (GW_n + GW_{n-1}), GW_k = W_{k+1} - W_k
  double aSuv = 0.0, aSw = 0.0;

  int s = 0;
  uint32_t phase = 0;
  auto step = [&](auto par_tag, const int n) {
    constexpr int C = decltype(par_tag)::value, Q = 1 - C;
    double (&uc)[RY] = cu[C], (&vc)[RY] = cv[C], (&wc)[RY] = cw[C], (&bc)[RY] = cb[C], (&pc)[RY] = cp[C];
    double (&up)[RY] = cu[Q], (&vp)[RY] = cv[Q], (&wp)[RY] = cw[Q], (&bp)[RY] = cb[Q], (&pp)[RY] = cp[Q];
    double (&nc)[RY] = cn[C], (&np)[RY] = cn[Q], (&XNc)[RY] = xn[C], (&XNp)[RY] = xn[Q], (&YNc)[RY] = yn[C], (&YNp)[RY] = yn[Q];
    mbar_wait(&full[s], phase);
    const TIN* T = tiles + s * (NF * TILE);
    const TIN* Tu = T, *Tv = T + TILE, *Tw = T + 2 * TILE, *Tb = T + F_B * TILE, *Tp = T + F_P * TILE, *Tn = T + F_N * TILE;
    if (++s == NS) { s = 0; phase ^= 1; }

    // ---- centre values of the thread's rows, the row to the south and the row to the north ----------------------------
#pragma unroll
    for (int r = 0; r < RY; ++r) {
      const int o = ro[r + 1];
      uc[r] = Tu[o]; vc[r] = Tv[o]; wc[r] = Tw[o];
      bc[r] = NEED_B ? Tb[o] : 0.0;
      pc[r] = NEED_P ? Tp[o] : 0.0;
      if (NUF) { nc[r] = Tn[o]; XNc[r] = Tn[o - 1] + nc[r]; YNc[r] = Tn[ro[r]] + nc[r]; }   // (every step: the next one's face means need them)
    }
    if (n > n_first) {
      const int oS = ro[0], oN = ro[RY + 1];
      const double uS = Tu[oS], wS = Tw[oS], vS = Tv[oS], vN = Tv[oN];
      const double bS = NEED_B ? Tb[oS] : 0.0, bN = (NEED_B && H_CDF) ? Tb[oN] : 0.0, pS = NEED_P ? Tp[oS] : 0.0;
      // plane classes (block-uniform): g0 = [this chunk owns step n and plane n is in the window], g1, g2: same for n-1, n-2
      const bool own = n >= k0;
      const bool in0 = n >= P.kz_lo && n <= P.kz_hi, in1 = (n - 1) >= P.kz_lo && (n - 1) <= P.kz_hi, in2 = (n - 2) >= P.kz_lo && (n - 2) <= P.kz_hi;
      const bool pfast = own && in0 && in1 && in2;
      const double g0 = (own && in0) ? 1.0 : 0.0, g1 = (own && in1) ? 1.0 : 0.0, g2 = (own && in2) ? 1.0 : 0.0;
      // z metrics of this step (STR): cell planes n, n-1, n-2 and faces n, n-1; G* = the window / chunk classes times dz_c
      double c0 = 1.0, c1 = 1.0, c2 = 1.0, dzf0 = 1.0, rzf0 = rdz, rzf1 = rdz, rzfn = rdz, rzc1 = rdz;
      if (STR) {
        const int lo = 1 - P.hz, hi = P.Nz + P.hz;
        const int q0 = min(max(n, lo), hi), q1 = min(max(n - 1, lo), hi), q2 = min(max(n - 2, lo), hi);
        c0 = __ldg(P.dzc + q0); c1 = __ldg(P.dzc + q1); c2 = __ldg(P.dzc + q2);
        dzf0 = __ldg(P.dzf + q0); rzf0 = __ldg(P.rdzf + q0); rzf1 = __ldg(P.rdzf + q1); rzc1 = __ldg(P.rdzc + q1);
        if (PZ) rzfn = __ldg(P.rdzf + min(max(n + 1, lo), hi));
      }
      const double G0 = g0 * c0, G1 = g1 * c1, G2 = g2 * c2;
      const double hzw = 0.5 * (c0 + c1);                               // FAST: volume weight of a z-face item
      const double rho0 = hzw * rzf0, rho1 = (0.5 * (c1 + c2)) * rzf1;  // (dz_c(n-1) + dz_c(n)) / (2 dz_f(n)): one for midpoint centres
      const double hzK = 0.5 * (G0 + G1), rho0g = hzK * rzf0, rho1g = (0.5 * (G1 + G2)) * rzf1;   // GEN: the same with the classes
      // mean-flow profiles around this plane (uniform loads; a missing profile is a ZeroField)
      double Un = 0.0, Un1 = 0.0, Vn = 0.0, Vn1 = 0.0, Wn = 0.0, dUz = 0.0, dVz = 0.0, dWz = 0.0;
      double GUc = 0.0, GUp = 0.0, GVc = 0.0, GVp = 0.0, GWc = 0.0, GWp = 0.0;
      if (PZ) {
        // (STR: the mean shear of a cell is the mean of its two face gradients, and it carries the cell's dz_c as volume weight)
        if (P.Uz) { const double a = __ldg(P.Uz + n + 1), d = __ldg(P.Uz + n - 2); Un = __ldg(P.Uz + n); Un1 = __ldg(P.Uz + n - 1);
                    dUz = Un - Un1;
                    if (!STR) { GUc = 0.5 * (a - Un1); GUp = 0.5 * (Un - d); }
                    else { GUc = (0.5 * c0) * fma(a - Un, rzfn, dUz * rzf0); GUp = (0.5 * c1) * fma(dUz, rzf0, (Un1 - d) * rzf1); } }
        if (P.Vz) { const double a = __ldg(P.Vz + n + 1), d = __ldg(P.Vz + n - 2); Vn = __ldg(P.Vz + n); Vn1 = __ldg(P.Vz + n - 1);
                    dVz = Vn - Vn1;
                    if (!STR) { GVc = 0.5 * (a - Vn1); GVp = 0.5 * (Vn - d); }
                    else { GVc = (0.5 * c0) * fma(a - Vn, rzfn, dVz * rzf0); GVp = (0.5 * c1) * fma(dVz, rzf0, (Vn1 - d) * rzf1); } }
        if (P.Wz) { const double a = __ldg(P.Wz + n + 1), d = __ldg(P.Wz + n - 1); Wn = __ldg(P.Wz + n); GWc = a - Wn; GWp = Wn - d; dWz = GWp; }
      }
      // one row of this thread at step n.  FAST (compile-time tag): every class weight is the interior one
      auto do_row = [&](auto fast_tag, const int r) {
        constexpr bool FAST = decltype(fast_tag)::value;
        const int o = ro[r + 1];
        const double us = r ? uc[r > 0 ? r - 1 : 0] : uS, ws = r ? wc[r > 0 ? r - 1 : 0] : wS;
        const double vn = (r < RY - 1) ? vc[r < RY - 1 ? r + 1 : 0] : vN;
        const double ue = Tu[o + 1], vw = Tv[o - 1], ww = Tw[o - 1];
        // undivided differences shared by the strain sums, the diagonal strains and the by-parts advection
        const double duy = uc[r] - us, dvx = vc[r] - vw, duz = uc[r] - up[r], dwx = wc[r] - ww, dvz = vc[r] - vp[r], dwy = wc[r] - ws;
        const double ax = ue - uc[r], ay = vn - vc[r], dwz = wc[r] - wp[r];
        double S12 = 0.0, S13 = 0.0, S23 = 0.0;
        if (NEED_S) {
          S12 = fma(dvx, rdx, duy * rdy);
          S13 = fma(dwx, rdx, duz * (STR ? rzf0 : rdz));
          S23 = fma(dwy, rdy, dvz * (STR ? rzf0 : rdz));
        }
        double P12 = 0.0, P13 = 0.0, P23 = 0.0, Uu = 0.0, Vv = 0.0, Ww = 0.0, P13w = 0.0, P23w = 0.0;
        if (H_ADV) {
          P12 = (vw + vc[r]) * (us + uc[r]);
          P13 = (ww + wc[r]) * (up[r] + uc[r]);
          P23 = (ws + wc[r]) * (vp[r] + vc[r]);
          if (STR) { P13w = (ww + wc[r]) * fma(c1, up[r], c0 * uc[r]); P23w = (ws + wc[r]) * fma(c1, vp[r], c0 * vc[r]); }
          const double su = uc[r] + ue, sv = vc[r] + vn, sw = wp[r] + wc[r];
          Uu = su * su; Vv = sv * sv; Ww = sw * sw;
        }
        const double ps = r ? pc[r > 0 ? r - 1 : 0] : pS;
        const double dpx = NEED_P ? pc[r] - Tp[o - 1] : 0.0, dpy = pc[r] - ps, dpz = pc[r] - pp[r];
        const double bs = r ? bc[r > 0 ? r - 1 : 0] : bS;
        const double dbx = NEED_B ? bc[r] - Tb[o - 1] : 0.0, dby = bc[r] - bs, dbz = bc[r] - bp[r];
        // NUF: the diffusivity of each face (kappa_const + kappa_scale x the two-point mean of nu_e: kappa_e = nu_e / Pr) rides on
        // the gradient of that face; otherwise kappa is one factor after the loop
        double dbxw = dbx, dbyw = dby, dbzw = dbz, kxe = 1.0, kyn = 1.0;
        if (NUF && (H_CHI || H_CDF)) {
          const double hs = 0.5 * P.kappa_scale;
          dbxw = fma(hs, XNc[r], P.kappa) * dbx; dbyw = fma(hs, YNc[r], P.kappa) * dby; dbzw = fma(hs, np[r] + nc[r], P.kappa) * dbz;
          if (H_CDF) { kxe = fma(hs, nc[r] + Tn[o + 1], P.kappa); kyn = fma(hs, nc[r] + Tn[ro[r + 2]], P.kappa); }
        }
        double lap = 0.0;
        if (H_CDF) {
          const double bn = (r < RY - 1) ? bc[r < RY - 1 ? r + 1 : 0] : bN;
          const double de = Tb[o + 1] - bc[r], dn = bn - bc[r];
          const double cdxy = NUF ? fma(dbxw - kxe * de, rdx2, (dbyw - kyn * dn) * rdy2) : fma(dbx - de, rdx2, (dby - dn) * rdy2);
          if (!STR) {
            lap = fma(-dbzw, rdz2, CDacc[r]);              // Laplacian (sign of div q) of the cell below, closed by this face
            CDacc[r] = fma(dbzw, rdz2, cdxy);
          } else {                                         // dz_c x the Laplacian: dz_c (in-plane part) + the two face gradients
            const double gl = dbzw * rzf0;
            lap = CDacc[r] - gl;
            CDacc[r] = fma(c0, cdxy, gl);
          }
        }
        // perturbations about the profiles: this plane's (u', v' at level n, w' at face n) and the previous plane's u', v'
        const double ucq = uc[r] - Un, vcq = vc[r] - Vn, wcq = wc[r] - Wn, upq = up[r] - Un1, vpq = vp[r] - Vn1;
        const double S13q = H_EPSP ? fma(-(STR ? rzf0 : rdz), dUz, S13) : 0.0, S23q = H_EPSP ? fma(-(STR ? rzf0 : rdz), dVz, S23) : 0.0, dwzq = dwz - dWz;
        // first factors of the dissipation's items: the item's viscosity times the strain (NUF; otherwise nu is applied after the loop)
        double s12w = S12, s13w = S13, s23w = S23, axw = ax, ayw = ay, dwzw = dwz, s13qw = S13q, s23qw = S23q, dwzqw = dwzq;
        if (NUF) {
          const double nffc = fma(0.25, (Tn[ro[r] - 1] + Tn[ro[r]]) + XNc[r], P.nu);
          const double nfcf = fma(0.25, XNp[r] + XNc[r], P.nu), ncff = fma(0.25, YNp[r] + YNc[r], P.nu);
          const double ncc = P.nu + nc[r], nccp = P.nu + np[r];
          s12w = nffc * S12; s13w = nfcf * S13; s23w = ncff * S23; axw = ncc * ax; ayw = ncc * ay; dwzw = nccp * dwz;
          if (H_EPSP) { s13qw = nfcf * S13q; s23qw = ncff * S23q; dwzqw = nccp * dwzq; }
        }
        if (FAST && STR) {
          if (H_EPSP) {
            aEeP = fma(c0 * s12w, S12, aEeP); aEeP = fma(dzf0 * s13qw, S13q, aEeP); aEeP = fma(dzf0 * s23qw, S23q, aEeP);
            aEzP = fma(rzc1 * dwzqw, dwzq, aEzP);
          }
          if (H_EPSP && !H_EPS) { aEx = fma(c0 * axw, ax, aEx); aEy = fma(c0 * ayw, ay, aEy); }
          if (H_TKE) { aKP = fma(c0 * ucq, ucq, aKP); aKP = fma(c0 * vcq, vcq, aKP); aKP = fma(hzw * wcq, wcq, aKP); }
          if (H_SP) {                                        // (the weights are in GU*, GV*; dz_c cancels in the W part)
            aSuv = fma((ww - Wn) + wcq, fma(GUc, ucq, GUp * upq), aSuv);
            aSuv = fma((ws - Wn) + wcq, fma(GVc, vcq, GVp * vpq), aSuv);
            aSw = fma(wcq * wcq, GWc + GWp, aSw);
          }
          if (H_EPS) {
            aEe = fma(c0 * s12w, S12, aEe); aEe = fma(dzf0 * s13w, S13, aEe); aEe = fma(dzf0 * s23w, S23, aEe);
            aEx = fma(c0 * axw, ax, aEx); aEy = fma(c0 * ayw, ay, aEy); aEz = fma(rzc1 * dwzw, dwz, aEz);
          }
          if (H_ADV) {
            aAe = fma(c0 * P12, S12, aAe);
            aAe = fma(P13w, (rho0 * rdx) * dwx, aAe); aAe = fma(P13, duz, aAe);
            aAe = fma(P23w, (rho0 * rdy) * dwy, aAe); aAe = fma(P23, dvz, aAe);
            aAx = fma(c0 * Uu, ax, aAx); aAy = fma(c0 * Vv, ay, aAy); aAz = fma(Ww, fma(rho0, wc[r], -(rho1 * wp[r])), aAz);
          }
          if (H_K) { aK = fma(c0 * uc[r], uc[r], aK); aK = fma(c0 * vc[r], vc[r], aK); aK = fma(hzw * wc[r], wc[r], aK); }
          if (H_PR) { aPx = fma(c0 * uc[r], dpx, aPx); aPy = fma(c0 * vc[r], dpy, aPy); aPz = fma(rho0 * wc[r], dpz, aPz); }
          if (H_WB) aW = fma(hzw * wc[r], bp[r] + bc[r], aW);
          if (H_CHI) { aCx = fma(c0 * dbxw, dbx, aCx); aCy = fma(c0 * dbyw, dby, aCy); aCz = fma(rzf0 * dbzw, dbz, aCz); }
          if (H_CDF) aD = fma(bp[r], lap, aD);
        } else if (FAST) {
          if (H_EPSP) {
            aEeP = fma(s12w, S12, aEeP); aEeP = fma(s13qw, S13q, aEeP); aEeP = fma(s23qw, S23q, aEeP);
            aEzP = fma(dwzqw, dwzq, aEzP);
          }
          if (H_EPSP && !H_EPS) { aEx = fma(axw, ax, aEx); aEy = fma(ayw, ay, aEy); }
          if (H_TKE) { aKP = fma(ucq, ucq, aKP); aKP = fma(vcq, vcq, aKP); aKP = fma(wcq, wcq, aKP); }
          if (H_SP) {
            aSuv = fma((ww - Wn) + wcq, fma(GUc, ucq, GUp * upq), aSuv);
            aSuv = fma((ws - Wn) + wcq, fma(GVc, vcq, GVp * vpq), aSuv);
            aSw = fma(wcq * wcq, GWc + GWp, aSw);
          }
          if (H_EPS) {
            aEe = fma(s12w, S12, aEe); aEe = fma(s13w, S13, aEe); aEe = fma(s23w, S23, aEe);
            aEx = fma(axw, ax, aEx); aEy = fma(ayw, ay, aEy); aEz = fma(dwzw, dwz, aEz);
          }
          if (H_ADV) {
            aAe = fma(P12, S12, aAe); aAe = fma(P13, S13, aAe); aAe = fma(P23, S23, aAe);
            aAx = fma(Uu, ax, aAx); aAy = fma(Vv, ay, aAy); aAz = fma(Ww, dwz, aAz);
          }
          if (H_K) { aK = fma(uc[r], uc[r], aK); aK = fma(vc[r], vc[r], aK); aK = fma(wc[r], wc[r], aK); }
          if (H_PR) { aPx = fma(uc[r], dpx, aPx); aPy = fma(vc[r], dpy, aPy); aPz = fma(wc[r], dpz, aPz); }
          if (H_WB) aW = fma(wc[r], bp[r] + bc[r], aW);
          if (H_CHI) { aCx = fma(dbxw, dbx, aCx); aCy = fma(dbyw, dby, aCy); aCz = fma(dbzw, dbz, aCz); }
          if (H_CDF) aD = fma(bp[r], lap, aD);
        } else if (lane_ok && STR) {
          // ---- GEN on a stretched axis: the class weights of the regular form below, times the metrics of their items ----
          const int j = j0 + jr0 + r;
          const double al = (ring || (j >= P.jy_lo && j <= P.jy_hi)) ? 1.0 : 0.0, be = (ring || (j - 1 >= P.jy_lo && j - 1 <= P.jy_hi)) ? 1.0 : 0.0,
                       an = (ring || (j + 1 >= P.jy_lo && j + 1 <= P.jy_hi)) ? 1.0 : 0.0;
          const double hy = 0.5 * (al + be), hyn = 0.5 * (al + an);
          const double hz = 0.5 * (g0 + g1);                               // cells sharing face n, halved
          if (H_EPS) {
            aEe = fma(hy * G0 * s12w, S12, aEe); aEe = fma(al * (hz * dzf0) * s13w, S13, aEe); aEe = fma(hy * (hz * dzf0) * s23w, S23, aEe);
            aEz = fma(al * g1 * rzc1 * dwzw, dwz, aEz);
          }
          if (H_EPS || H_EPSP) { aEx = fma(al * G0 * axw, ax, aEx); aEy = fma(al * G0 * ayw, ay, aEy); }
          if (H_EPSP) {
            aEeP = fma(hy * G0 * s12w, S12, aEeP); aEeP = fma(al * (hz * dzf0) * s13qw, S13q, aEeP); aEeP = fma(hy * (hz * dzf0) * s23qw, S23q, aEeP);
            aEzP = fma(al * g1 * rzc1 * dwzqw, dwzq, aEzP);
          }
          if (H_TKE) { aKP = fma(al * G0 * ucq, ucq, aKP); aKP = fma(hy * G0 * vcq, vcq, aKP); aKP = fma(al * hzK * wcq, wcq, aKP); }
          if (H_SP) {
            aSuv = fma(al * ((ww - Wn) + wcq), fma(g0 * GUc, ucq, g1 * GUp * upq), aSuv);
            aSuv = fma(fma(be, ws - Wn, al * wcq), fma(g0 * GVc, vcq, g1 * GVp * vpq), aSuv);
            aSw = fma(al * wcq * wcq, fma(g0, GWc, g1 * GWp), aSw);
          }
          if (H_ADV) {
            aAe = fma(G0 * P12, fma(rdx * hy, dvx, rdy * (al * uc[r] - be * us)), aAe);
            aAe = fma(al * P13w, (rdx * rho0g) * dwx, aAe); aAe = fma(al * P13, g0 * uc[r] - g1 * up[r], aAe);
            aAe = fma(P23w, (rdy * rho0g) * (al * wc[r] - be * ws), aAe); aAe = fma(hy * P23, g0 * vc[r] - g1 * vp[r], aAe);
            aAx = fma(al * G0 * Uu, ax, aAx);
            aAy = fma(G0 * Vv, hyn * vn - hy * vc[r], aAy);
            if (!ring && j == P.jy_lo) {
              const double vs = r ? vc[r > 0 ? r - 1 : 0] : vS;
              aAy = fma(G0 * ((vs + vc[r]) * (vs + vc[r])), 0.5 * vc[r], aAy);
            }
            aAz = fma(al * Ww, rho0g * wc[r] - rho1g * wp[r], aAz);
          }
          if (H_K) { aK = fma(al * G0 * uc[r], uc[r], aK); aK = fma(hy * G0 * vc[r], vc[r], aK); aK = fma(al * hzK * wc[r], wc[r], aK); }
          if (H_PR) { aPx = fma(al * G0 * uc[r], dpx, aPx); aPy = fma(hy * G0 * vc[r], dpy, aPy); aPz = fma(al * rho0g * wc[r], dpz, aPz); }
          if (H_WB) aW = fma(al * hzK * wc[r], bp[r] + bc[r], aW);
          if (H_CHI) { aCx = fma(al * G0 * dbxw, dbx, aCx); aCy = fma(hy * G0 * dbyw, dby, aCy); aCz = fma(al * (hz * rzf0) * dbzw, dbz, aCz); }
          if (H_CDF) aD = fma(al * g1 * bp[r], lap, aD);
        } else if (lane_ok) {
          // ---- GEN: a window-edge row or plane (or a partly filled x tile): explicit class weights (all one elsewhere) ----
          const int j = j0 + jr0 + r;
          const double al = (ring || (j >= P.jy_lo && j <= P.jy_hi)) ? 1.0 : 0.0, be = (ring || (j - 1 >= P.jy_lo && j - 1 <= P.jy_hi)) ? 1.0 : 0.0,
                       an = (ring || (j + 1 >= P.jy_lo && j + 1 <= P.jy_hi)) ? 1.0 : 0.0;
          const double hy = 0.5 * (al + be), hyn = 0.5 * (al + an);        // y faces / corner rows j and j+1: half their cell count
          const double hz = 0.5 * (g0 + g1), hzp = 0.5 * (g1 + g2);        // z faces n and n-1
          if (H_EPS) {
            aEe = fma(hy * g0 * s12w, S12, aEe); aEe = fma(al * hz * s13w, S13, aEe); aEe = fma(hy * hz * s23w, S23, aEe);
            aEz = fma(al * g1 * dwzw, dwz, aEz);
          }
          if (H_EPS || H_EPSP) { aEx = fma(al * g0 * axw, ax, aEx); aEy = fma(al * g0 * ayw, ay, aEy); }
          if (H_EPSP) {
            aEeP = fma(hy * g0 * s12w, S12, aEeP); aEeP = fma(al * hz * s13qw, S13q, aEeP); aEeP = fma(hy * hz * s23qw, S23q, aEeP);
            aEzP = fma(al * g1 * dwzqw, dwzq, aEzP);
          }
          if (H_TKE) { aKP = fma(al * g0 * ucq, ucq, aKP); aKP = fma(hy * g0 * vcq, vcq, aKP); aKP = fma(al * hz * wcq, wcq, aKP); }
          if (H_SP) {
            // x faces belong to one row; a y face j to the cells j-1 (weight be) and j (weight al); the part with GU_n to the
            // cells of plane n (g0), the part with GU_{n-1} to those of plane n-1 (g1)
            aSuv = fma(al * ((ww - Wn) + wcq), fma(g0 * GUc, ucq, g1 * GUp * upq), aSuv);
            aSuv = fma(fma(be, ws - Wn, al * wcq), fma(g0 * GVc, vcq, g1 * GVp * vpq), aSuv);
            aSw = fma(al * wcq * wcq, fma(g0, GWc, g1 * GWp), aSw);
          }
          if (H_ADV) {
            // the by-parts weights of a window edge: the boundary face keeps the momentum on the window's side only
            aAe = fma(g0 * P12, fma(rdx * hy, dvx, rdy * (al * uc[r] - be * us)), aAe);
            aAe = fma(al * P13, fma(rdx * hz, dwx, rdz * (g0 * uc[r] - g1 * up[r])), aAe);
            aAe = fma(P23, fma(rdy * hz, al * wc[r] - be * ws, rdz * hy * (g0 * vc[r] - g1 * vp[r])), aAe);
            aAx = fma(al * g0 * Uu, ax, aAx);
            aAy = fma(g0 * Vv, hyn * vn - hy * vc[r], aAy);
            if (!ring && j == P.jy_lo) {                   // the flux cell south of the first window row, seen from face jy_lo only
              const double vs = r ? vc[r > 0 ? r - 1 : 0] : vS;
              aAy = fma(g0 * ((vs + vc[r]) * (vs + vc[r])), 0.5 * vc[r], aAy);
            }
            aAz = fma(al * Ww, hz * wc[r] - hzp * wp[r], aAz);
          }
          if (H_K) { aK = fma(al * g0 * uc[r], uc[r], aK); aK = fma(hy * g0 * vc[r], vc[r], aK); aK = fma(al * hz * wc[r], wc[r], aK); }
          if (H_PR) { aPx = fma(al * g0 * uc[r], dpx, aPx); aPy = fma(hy * g0 * vc[r], dpy, aPy); aPz = fma(al * hz * wc[r], dpz, aPz); }
          if (H_WB) aW = fma(al * hz * wc[r], bp[r] + bc[r], aW);
          if (H_CHI) { aCx = fma(al * g0 * dbxw, dbx, aCx); aCy = fma(hy * g0 * dbyw, dby, aCy); aCz = fma(al * hz * dbzw, dbz, aCz); }
          if (H_CDF) aD = fma(al * g1 * bp[r], lap, aD);
        }
      };
      if (pfast && tfast) {
        // straight-line code over the thread's rows: the compiler interleaves their independent chains
#pragma unroll
        for (int r = 0; r < RY; ++r) do_row(std::true_type{}, r);
      } else {
#pragma unroll
        for (int r = 0; r < RY; ++r) {
          if (rlive[r]) do_row(std::false_type{}, r);
        }
      }
    }
    __syncthreads();
    if (tid == 0 && n + NS <= n_last) issue(n + NS);       // every thread is done with this ring slot: refill it
  };
  for (int n = n_first;;) {
    step(std::integral_constant<int, 0>{}, n);
    if (++n > n_last) break;
    step(std::integral_constant<int, 1>{}, n);
    if (++n > n_last) break;
  }

  // ---- class sums -> the seven integrals (per-cell scale factors of the reference formulas folded in) ---------------
  double acc[FD_N];
#pragma unroll
  for (int d = 0; d < FD_N; ++d) acc[d] = 0.0;
  const double fz = STR ? 1.0 : rdz, fz2 = STR ? 1.0 : rdz2;          // STR: the z metrics are inside the sums
  acc[FD_K] = 0.5 * aK;
  acc[FD_ADV] = -0.25 * (aAe + (rdx * aAx + rdy * aAy + fz * aAz));
  acc[FD_PR] = rdx * aPx + rdy * aPy + fz * aPz;
  acc[FD_WB] = 0.5 * P.ghat_z * aW;
  acc[FD_EPS] = (NUF ? 1.0 : P.nu) * (aEe + 2.0 * (rdx2 * aEx + rdy2 * aEy + fz2 * aEz));      // NUF: the viscosities are inside the sums
  acc[FD_CHI] = 2.0 * (NUF ? 1.0 : P.kappa) * (rdx2 * aCx + rdy2 * aCy + fz2 * aCz);
  acc[FD_CDF] = 2.0 * (NUF ? 1.0 : P.kappa) * aD;
  acc[FD_EPSP] = (NUF ? 1.0 : P.nu) * (aEeP + 2.0 * (rdx2 * aEx + rdy2 * aEy + fz2 * aEzP));
  acc[FD_TKE] = 0.5 * aKP;
  acc[FD_SP] = -fz * (0.25 * aSuv + 0.5 * aSw);
  if (P.nslots > 0) {
    const unsigned FULLMASK = 0xffffffffu;
#pragma unroll
    for (int d = 0; d < FD_N; ++d) {
      if (!(MASK & (1 << d)) || P.o[d].slot < 0) continue;
      double v = acc[d];
      for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(FULLMASK, v, o);
      __syncthreads();
      if (lane == 0) red[warp] = v;
      __syncthreads();
      if (tid == 0) {
        double t = 0.0;
        for (int w = 0; w < NW; ++w) t += red[w];
        // a halo wait that gave up poisons this block's partials: the (all-reduced) integrals read NaN on every rank
        P.partial[(size_t)blockIdx.x * P.nslots + P.o[d].slot] = gate_timed_out ? __longlong_as_double(0x7ff8000000000000LL) : t * P.V;
      }
    }
  }
}
