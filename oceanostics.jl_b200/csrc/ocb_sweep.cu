// K1: the fused multi-diagnostic stencil sweep -- plan object, the generic kernel (any mix of the
// OCB_DIAG_* requests, any operand kind, stretched z, Float32/Float64) and the dispatch to the
// register-pipelined budget kernel (ocb_sweep_fused.cu) when a plan matches its envelope.
//
// Replaces, per plan, what the reference does with one Oceananigans launch per KernelFunctionOperation
// (`launch!(arch, grid, params, _compute!, data, op)`) plus one `sum!` sweep per Integral
// (the explicit form of that recipe is src/SpatialFilters/SpatialFilters.jl:325-330).
#include "ocb_sweep_plan.cuh"
#include "ocb_sweep_fused.cuh"
#include "ocb_sweep_strain.cuh"
#include <new>

// ---- z metric tables -------------------------------------------------------------------------------------
// table entry m <-> k = m + 1 - Hz, m = 0 .. Nz + 2Hz  (Nz + 2Hz + 1 entries per table)
template <class T>
__global__ void ocb_ztables_kernel(int n, int Hz, T dz_uniform, T z_bottom, const T* __restrict__ dzc_in,
                                   const T* __restrict__ dzf_in, const T* __restrict__ zc_in, const T* __restrict__ zf_in,
                                   T* __restrict__ tab) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= n) return;
  const int k = m + 1 - Hz;
  T dzc, dzf, zc, zf;
  if (dzc_in) {
    const int mc = m < n - 1 ? m : n - 2;   // Δzᵃᵃᶜ / zᵃᵃᶜ have one entry fewer: repeat the last
    dzc = dzc_in[mc]; zc = zc_in ? zc_in[mc] : T(0);
    dzf = dzf_in[m];  zf = zf_in ? zf_in[m] : T(0);
  } else {
    dzc = dzf = dz_uniform;
    zf = z_bottom + dz_uniform * T(k - 1);
    zc = z_bottom + dz_uniform * (T(k) - T(0.5));
  }
  tab[0 * n + m] = dzc; tab[1 * n + m] = dzf;
  tab[2 * n + m] = T(1) / dzc; tab[3 * n + m] = T(1) / dzf;
  tab[4 * n + m] = zc; tab[5 * n + m] = zf;
}

// ---- generic kernel --------------------------------------------------------------------------------------
template <class T, int DIAG, bool PERT>
__device__ __forceinline__ void ocb_run_request(const SweepParams<T>& P, const DReq& r, int i, int j, int k0, int k1,
                                                double* sh, int blk) {
  const bool in_ij = i >= r.lo[0] && i <= r.hi[0] && j >= r.lo[1] && j <= r.hi[1];
  const int ka = k0 > r.lo[2] ? k0 : r.lo[2];
  const int kb = k1 < r.hi[2] ? k1 : r.hi[2];
  double acc = 0.0;
  if (in_ij && ka <= kb) {
    // thread-local views shifted to the first cell of this column; the evaluators then run at the literal origin
    SweepIn<T> L = P.in;
    DGrid<T> G = P.g;
    ocb_shift_inputs(L, G, i, j, ka);
    const Ctx<T> X(G, L, PERT);
    T* out = (T*)r.out;
    if (out) out += i * r.os[0] + j * r.os[1] + (long long)ka * r.os[2];
    const bool integrate = r.slot >= 0;
    const bool zface = r.loc[2] == OCB_FACE;
    const double dxdy = (double)P.g.dx * (double)P.g.dy;
    for (int k = ka; k <= kb; ++k) {
      const T val = Eval<T, DIAG>::at(X, r.flags, 0, 0, 0);
      if (out) { *out = val; out += r.os[2]; }
      if (integrate) acc += (double)val * (dxdy * (double)(zface ? G.dzf[0] : G.dzc[0]));
      ocb_shift_inputs(L, G, 0, 0, 1);
    }
  }
  if (r.slot >= 0) {   // uniform across the block
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    __syncthreads();
    if (lane == 0) sh[warp] = acc;
    __syncthreads();
    if (tid == 0) {
      double s = 0.0;
      const int nw = (blockDim.x * blockDim.y) >> 5;
      for (int w = 0; w < nw; ++w) s += sh[w];
      P.partial[(size_t)blk * P.nslots + r.slot] = s;
    }
  }
}

template <class T, int DIAG>
__device__ __forceinline__ void ocb_run_request_dispatch(const SweepParams<T>& P, const DReq& r, int i, int j, int k0, int k1,
                                                         double* sh, int blk) {
  if constexpr (ocb_diag_takes_perturbation(DIAG)) {
    if (r.flags & OCB_REQ_PERTURBATION) { ocb_run_request<T, DIAG, true>(P, r, i, j, k0, k1, sh, blk); return; }
  }
  ocb_run_request<T, DIAG, false>(P, r, i, j, k0, k1, sh, blk);
}

template <class T>
__global__ void __launch_bounds__(256, 3) ocb_sweep_generic_kernel(const __grid_constant__ SweepParams<T> P) {
  __shared__ double sh[8];
  const int i = P.lo[0] + blockIdx.x * blockDim.x + threadIdx.x;
  const int j = P.lo[1] + blockIdx.y * blockDim.y + threadIdx.y;
  const int k0 = P.lo[2] + blockIdx.z * P.kchunk;
  const int k1 = min(k0 + P.kchunk - 1, P.hi[2]);
  const int blk = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
  for (int r = 0; r < P.nreq; ++r) {
    const DReq& R = P.req[r];
    switch (R.diag) {
#define OCB_CASE(D) case D: ocb_run_request_dispatch<T, D>(P, R, i, j, k0, k1, sh, blk); break;
      OCB_FOR_EACH_DIAG(OCB_CASE)
#undef OCB_CASE
      default: break;
    }
  }
}

// ---- plan ------------------------------------------------------------------------------------------------
struct ocb_plan {
  int dtype;
  int variant;            // which kernels one execute launches: 0 generic, 1 budget, 2 budget + generic,
                          // 3 velocity-gradient, 4 velocity-gradient + generic, 5 budget + velocity-gradient, 6 all three
  SweepParams<double> p64;
  SweepParams<float> p32;
  void* ztab;             // library-owned z metric tables
  double* partial;        // per-block partial integrals
  dim3 grid, block;
  int nblocks, fused_blocks, strain_blocks, generic_blocks, nslots;
  int n_launches;
  double algorithmic_bytes;
  int device;             // the device that owns the operands: every call of this plan runs there
  ocb_fused_plan* fused;
  ocb_strain_plan* strain;
};

template <class T>
static int build_ztables(const ocb_grid* grid, void** ztab_out, ZTables<T>* zt) {
  const int Hz = grid->H[2], n = grid->N[2] + 2 * Hz + 1;
  T* tab = nullptr;
  OCB_CUDA(cudaMalloc(&tab, sizeof(T) * 6 * n));
  if (grid->dzc || grid->dzf) OCB_REQUIRE(grid->dzc && grid->dzf, "a stretched z axis needs both dzc and dzf");
  ocb_ztables_kernel<T><<<(n + 127) / 128, 128>>>(n, Hz, T(grid->d[2]), T(grid->z_bottom), (const T*)grid->dzc, (const T*)grid->dzf,
                                                 (const T*)grid->zc, (const T*)grid->zf, tab);
  OCB_CUDA(cudaGetLastError());
  const int off = Hz - 1;   // table[k] = tab[k + Hz - 1]
  zt->dzc = tab + 0 * n + off; zt->dzf = tab + 1 * n + off; zt->rdzc = tab + 2 * n + off; zt->rdzf = tab + 3 * n + off;
  zt->zc = tab + 4 * n + off;  zt->zf = tab + 5 * n + off;
  *ztab_out = tab;
  return OCB_OK;
}

static double count_algorithmic_bytes(const ocb_grid* g, const ocb_sweep_inputs* in, const ocb_request* reqs, int n) {
  // (distinct full-size input fields read + output fields written) * sizeof(FT) * Nx*Ny*Nz
  bool use[16] = {false};   // u v w b p U V W B aux0..6
  auto mark_uvw = [&]() { use[0] = use[1] = use[2] = true; };
  for (int r = 0; r < n; ++r) {
    const int d = reqs[r].diag;
    const bool pert = reqs[r].flags & OCB_REQ_PERTURBATION;
    switch (d) {
      case OCB_DIAG_PR: mark_uvw(); use[4] = true; break;
      case OCB_DIAG_WB: mark_uvw(); use[3] = true; break;
      case OCB_DIAG_CHI: case OCB_DIAG_CDIFF: use[3] = true; if (pert) use[8] = true; break;
      case OCB_DIAG_RI: case OCB_DIAG_EPV: case OCB_DIAG_DEPV: case OCB_DIAG_TWPV: mark_uvw(); use[3] = true; break;
      case OCB_DIAG_BPE: use[3] = true; use[15] = true; break;
      case OCB_DIAG_UPSILON: use[15] = true; break;
      case OCB_DIAG_APE_EPS: use[3] = true; use[13] = true; break;
      case OCB_DIAG_APE: use[3] = true; use[14] = use[15] = true; break;
      case OCB_DIAG_FEPS: case OCB_DIAG_PI: mark_uvw(); for (int a = 0; a < 6; ++a) use[9 + a] = true; break;
      case OCB_DIAG_SFEPS: mark_uvw(); for (int a = 0; a < 7; ++a) use[9 + a] = true; break;
      case OCB_DIAG_SFKE: mark_uvw(); use[15] = true; break;
      case OCB_DIAG_KE_FORCING: mark_uvw(); use[9] = use[10] = use[11] = true; break;
      case OCB_DIAG_C_TENDENCY: mark_uvw(); use[3] = true; use[9] = true; break;
      case OCB_DIAG_KE_TENDENCY: mark_uvw(); use[3] = use[4] = true; use[9] = use[10] = use[11] = true; break;
      case OCB_DIAG_NU_SMAG: mark_uvw(); if (in->les[1] != 0) use[3] = true; break;
      case OCB_DIAG_U_TERM: case OCB_DIAG_V_TERM: case OCB_DIAG_W_TERM: mark_uvw(); use[3] = use[4] = true; use[9] = use[10] = use[11] = true; break;
      case OCB_DIAG_C_TERM: mark_uvw(); use[3] = true; use[9] = true; break;
      case OCB_DIAG_QX: case OCB_DIAG_QY: case OCB_DIAG_QZ: use[3] = true; break;
      default: mark_uvw(); break;
    }
    if (pert || d == OCB_DIAG_TKE || (d >= OCB_DIAG_SPX && d <= OCB_DIAG_SP)) use[5] = use[6] = use[7] = true;
  }
  const ocb_field* f[16] = {&in->u, &in->v, &in->w, &in->b, &in->p, &in->U, &in->V, &in->W, &in->B,
                            &in->aux[0], &in->aux[1], &in->aux[2], &in->aux[3], &in->aux[4], &in->aux[5], &in->aux[6]};
  int nin = 0;
  for (int a = 0; a < 16; ++a) {
    if (!use[a] || f[a]->kind != OCB_FIELD_ARRAY || !f[a]->data) continue;
    bool full = true;
    for (int d = 0; d < 3; ++d) if (f[a]->loc[d] == OCB_NOTHING) full = false;
    if (full) ++nin;
  }
  bool eddy = false;
  for (int r = 0; r < n; ++r) {
    const int d = reqs[r].diag;
    if (d == OCB_DIAG_EPS || d == OCB_DIAG_EPS_ISO || d == OCB_DIAG_CHI || d == OCB_DIAG_CDIFF || d == OCB_DIAG_APE_EPS ||
        d == OCB_DIAG_KE_STRESS || d == OCB_DIAG_C_TENDENCY || d == OCB_DIAG_KE_TENDENCY || (d >= OCB_DIAG_U_TERM && d <= OCB_DIAG_QZ) || (d >= OCB_DIAG_VF11 && d <= OCB_DIAG_VF23)) eddy = true;
  }
  if (eddy) {   // count each distinct eddy field once
    const void* seen[2 * OCB_MAX_CLOSURE_FIELDS]; int ns = 0;
    auto add = [&](const ocb_field& fld) { for (int s = 0; s < ns; ++s) if (seen[s] == fld.data) return; seen[ns++] = fld.data; ++nin; };
    for (int a = 0; a < in->closure.n_nu_fields; ++a) add(in->closure.nu_fields[a]);
    for (int a = 0; a < in->closure.n_kappa_fields; ++a) add(in->closure.kappa_fields[a]);
  }
  int nout = 0;
  for (int r = 0; r < n; ++r) if (reqs[r].out) ++nout;
  return (double)(nin + nout) * (double)ocb_sizeof(g->dtype) * (double)g->N[0] * (double)g->N[1] * (double)g->N[2];
}

// launch geometry of the generic kernel for the requests in *P
template <class T>
static void generic_geometry(ocb_plan* pl, SweepParams<T>* P) {
  const int ni = P->hi[0] - P->lo[0] + 1, nj = P->hi[1] - P->lo[1] + 1, nk = P->hi[2] - P->lo[2] + 1;
  pl->block = dim3(32, 8, 1);
  const int gx = (ni + 31) / 32, gy = (nj + 7) / 8;
  const int target = ocb_sm_count() * 16;
  int gz = (target + gx * gy - 1) / (gx * gy);
  if (gz > nk) gz = nk;
  if (gz < 1) gz = 1;
  int kchunk = (nk + gz - 1) / gz;
  if (kchunk > 128) kchunk = 128;
  gz = (nk + kchunk - 1) / kchunk;
  P->kchunk = kchunk;
  pl->grid = dim3(gx, gy, gz);
  pl->generic_blocks = gx * gy * gz;
}

// The planner: every request the register-pipelined budget kernel can evaluate goes to it (one launch), the rest to
// the generic kernel (one launch); both write per-block partials of the same integral vector.
//   variant 0: generic only, 1: budget kernel only, 2: split between the two.
template <class T>
static int finish_plan(ocb_plan* pl, const ocb_grid* grid, const ocb_sweep_inputs* in, const ocb_request* reqs, int n,
                       SweepParams<T>* P) {
  ZTables<T> zt;
  int rc = build_ztables<T>(grid, &pl->ztab, &zt);
  if (rc != OCB_OK) return rc;
  rc = ocb_build_sweep_params<T>(grid, in, reqs, n, zt, P);      // validates every request, numbers the slots
  if (rc != OCB_OK) return rc;
  const int nslots = P->nslots;
  pl->fused = nullptr;
  pl->variant = 0;
  // ---- fused subset ---------------------------------------------------------------------------------------
  ocb_request fr[OCB_MAX_REQ], gr[OCB_MAX_REQ];
  int nf = 0, ng = 0;
  // the dissipation rate is known to both pipelined kernels: next to other budget terms it stays with the budget
  // kernel, next to strain / rotation moduli only it goes with them (one sweep instead of two)
  int budget_other = 0, strain_other = 0;
  for (int r = 0; r < n; ++r) {
    if (reqs[r].diag == OCB_DIAG_EPS) continue;
    if (ocb_fused_handles_diag(reqs[r].diag) && !(reqs[r].flags & OCB_REQ_CENTERED4)) ++budget_other;
    if (ocb_strain_handles_diag(reqs[r].diag)) ++strain_other;
  }
  const bool eps_with_moduli = budget_other == 0 && strain_other > 0;
  for (int r = 0; r < n; ++r) {
    // (the pipelined kernels evaluate Centered second-order advection; a fourth-order request stays with the generic kernel)
    if (ocb_fused_handles_diag(reqs[r].diag) && !(reqs[r].flags & OCB_REQ_CENTERED4) && !(eps_with_moduli && reqs[r].diag == OCB_DIAG_EPS)) fr[nf++] = reqs[r];
    else gr[ng++] = reqs[r];
  }
  const bool stretched_z = grid->dzc || grid->dzf;
  if (in->closure.n_nu_fields || in->closure.n_kappa_fields) {
    // eddy-viscosity / diffusivity fields.  The budget kernels keep the closure-independent terms (K, ADV, PR, w b) and, as class
    // sums, the dissipation with one nu_e field; what they decline goes to the velocity-gradient kernel (dissipation) and the
    // generic one (tracer-variance terms).  Tried from the largest set down: everything, without the tracer-variance terms,
    // without the dissipation too.
    for (int stage = 0; stage < 3 && nf > 0 && !pl->fused; ++stage) {
      if (stage >= 1) {
        int kept = 0, moved = 0;
        for (int r = 0; r < nf; ++r) {
          const int dg = fr[r].diag;
          const bool go = stage == 1 ? (dg == OCB_DIAG_CHI || dg == OCB_DIAG_CDIFF) : dg == OCB_DIAG_EPS;
          if (go) { gr[ng++] = fr[r]; ++moved; } else fr[kept++] = fr[r];
        }
        nf = kept;
        if (!moved || nf == 0) continue;
      }
      SweepParams<T> Pf;
      rc = ocb_build_sweep_params<T>(grid, in, fr, nf, zt, &Pf);
      if (rc != OCB_OK) return rc;
      Pf.nslots = nslots;
      rc = ocb_fused_try_create(grid, in, fr, nf, Pf, &pl->fused);
      if (rc != OCB_OK) return rc;
    }
  }
  for (int attempt = 0; attempt < 4 && nf > 0 && !pl->fused; ++attempt) {
    if (attempt >= 1) {
      // the budget kernels declined the set: retry without the term with the narrowest envelope -- first the potential
      // vorticity, then the turbulent kinetic energy (class-sum kernel only), then (stretched z: the per-plane-metric variant
      // evaluates the seven plain budget terms) everything about mean-flow operands
      if (attempt == 3 && !stretched_z) break;
      const int drop = attempt == 1 ? OCB_DIAG_EPV : OCB_DIAG_TKE;
      int kept = 0, moved = 0;
      for (int r = 0; r < nf; ++r) {
        const bool go = attempt < 3 ? fr[r].diag == drop
                                    : ((fr[r].flags & OCB_REQ_PERTURBATION) || fr[r].diag == OCB_DIAG_SP || fr[r].diag == OCB_DIAG_SPZ);
        if (go) { gr[ng++] = fr[r]; ++moved; } else fr[kept++] = fr[r];
      }
      nf = kept;
      if (!moved) continue;
      if (nf == 0) break;
    }
    SweepParams<T> Pf;
    rc = ocb_build_sweep_params<T>(grid, in, fr, nf, zt, &Pf);
    if (rc != OCB_OK) return rc;
    Pf.nslots = nslots;
    rc = ocb_fused_try_create(grid, in, fr, nf, Pf, &pl->fused);
    if (rc != OCB_OK) return rc;
  }
  if (!pl->fused) {                                  // nothing went to the budget kernel: start over with every request
    ng = 0;
    for (int r = 0; r < n; ++r) gr[ng++] = reqs[r];
  }
  // ---- velocity-gradient subset of what is left (closure-dependent dissipation, strain / rotation moduli) -------------
  pl->strain = nullptr;
  {
    ocb_request sr[OCB_MAX_REQ], rest[OCB_MAX_REQ];
    int ns = 0, nrest = 0;
    for (int r = 0; r < ng; ++r) {
      if (ocb_strain_handles_diag(gr[r].diag)) sr[ns++] = gr[r]; else rest[nrest++] = gr[r];
    }
    if (ns > 0) {
      SweepParams<T> Ps;
      rc = ocb_build_sweep_params<T>(grid, in, sr, ns, zt, &Ps);
      if (rc != OCB_OK) return rc;
      Ps.nslots = nslots;
      rc = ocb_strain_try_create(grid, in, sr, ns, Ps, &pl->strain);
      if (rc != OCB_OK) return rc;
      if (pl->strain) {
        ng = nrest;
        for (int r = 0; r < nrest; ++r) gr[r] = rest[r];
      }
    }
  }
  const int fused_blocks = pl->fused ? ocb_fused_nblocks(pl->fused) : 0;
  const int strain_blocks = pl->strain ? ocb_strain_nblocks(pl->strain) : 0;
  if (pl->fused && pl->strain) pl->variant = ng > 0 ? 6 : 5;
  else if (pl->strain) pl->variant = ng > 0 ? 4 : 3;
  else if (pl->fused) pl->variant = ng > 0 ? 2 : 1;
  else pl->variant = 0;
  if ((pl->fused || pl->strain) && ng > 0) {        // the generic kernel takes only what is left
    rc = ocb_build_sweep_params<T>(grid, in, gr, ng, zt, P);
    if (rc != OCB_OK) return rc;
    P->nslots = nslots;
  }
  pl->generic_blocks = 0;
  if (ng > 0) generic_geometry<T>(pl, P);
  pl->nblocks = fused_blocks + strain_blocks + pl->generic_blocks;
  pl->fused_blocks = fused_blocks;
  pl->strain_blocks = strain_blocks;
  pl->nslots = nslots;
  pl->n_launches = (pl->fused ? 1 : 0) + (pl->strain ? 1 : 0) + (ng > 0 ? 1 : 0) + (nslots > 0 ? 1 : 0);
  if (nslots > 0) {
    // rows [0, fused_blocks) belong to the budget kernel, the next strain_blocks to the velocity-gradient kernel, the
    // rest to the generic kernel; a slot a kernel does not own
    // stays zero in its rows (the buffer is cleared once, here)
    OCB_CUDA(cudaMalloc(&pl->partial, sizeof(double) * (size_t)pl->nblocks * nslots));
    OCB_CUDA(cudaMemset(pl->partial, 0, sizeof(double) * (size_t)pl->nblocks * nslots));
    P->partial = pl->partial + (size_t)(fused_blocks + strain_blocks) * nslots;
  }
  OCB_CUDA(cudaDeviceSynchronize());   // plan creation (not execution) may block: tables are ready afterwards
  return OCB_OK;
}

extern "C" int ocb_plan_create(const ocb_grid* grid, const ocb_sweep_inputs* inputs, const ocb_request* requests,
                               int32_t n_requests, ocb_plan** plan_out) {
  OCB_REQUIRE(grid && inputs && requests && plan_out, "ocb_plan_create: null argument");
  OCB_REQUIRE(grid->dtype == OCB_F32 || grid->dtype == OCB_F64, "grid dtype must be OCB_F32 or OCB_F64");
  int rc = ocb_require_device();
  if (rc != OCB_OK) return rc;
  // the device of the first array operand (or output) is the plan's device
  const void* owner = nullptr;
  const ocb_field* ops[] = {&inputs->u, &inputs->v, &inputs->w, &inputs->b, &inputs->p, &inputs->U, &inputs->V, &inputs->W, &inputs->B};
  for (const ocb_field* f : ops) if (!owner && f->kind == OCB_FIELD_ARRAY && f->data) owner = f->data;
  for (int a = 0; a < OCB_N_AUX && !owner; ++a) if (inputs->aux[a].kind == OCB_FIELD_ARRAY && inputs->aux[a].data) owner = inputs->aux[a].data;
  for (int r = 0; r < n_requests && !owner; ++r) owner = requests[r].out;
  OcbDeviceGuard device_guard(owner);
  ocb_plan* pl = new (std::nothrow) ocb_plan();
  if (!pl) OCB_FAIL(OCB_ERR_ALLOC, "out of host memory");
  memset((void*)pl, 0, sizeof(*pl));
  OCB_CUDA(cudaGetDevice(&pl->device));
  pl->dtype = grid->dtype;
  pl->algorithmic_bytes = count_algorithmic_bytes(grid, inputs, requests, n_requests);
  rc = (grid->dtype == OCB_F64) ? finish_plan<double>(pl, grid, inputs, requests, n_requests, &pl->p64)
                                : finish_plan<float>(pl, grid, inputs, requests, n_requests, &pl->p32);
  if (rc != OCB_OK) { ocb_plan_destroy(pl); return rc; }
  *plan_out = pl;
  return OCB_OK;
}

extern "C" int ocb_plan_execute(ocb_plan* pl, double* integrals, void* stream) {
  OCB_REQUIRE(pl, "ocb_plan_execute: null plan");
  OcbDeviceGuard device_guard(pl->device);
  cudaStream_t st = (cudaStream_t)stream;
  const int nslots = pl->nslots;
  OCB_REQUIRE(nslots == 0 || integrals, "this plan integrates %d slot(s): `integrals` must be a device double[%d]", nslots, nslots);
  if (pl->fused) {
    int rc = ocb_fused_execute(pl->fused, pl->partial, st);
    if (rc != OCB_OK) return rc;
  }
  if (pl->strain) {
    int rc = ocb_strain_execute(pl->strain, pl->partial ? pl->partial + (size_t)pl->fused_blocks * nslots : nullptr, st);
    if (rc != OCB_OK) return rc;
  }
  if (pl->generic_blocks > 0) {
    if (pl->dtype == OCB_F64) ocb_sweep_generic_kernel<double><<<pl->grid, pl->block, 0, st>>>(pl->p64);
    else ocb_sweep_generic_kernel<float><<<pl->grid, pl->block, 0, st>>>(pl->p32);
  }
  OCB_CUDA(cudaGetLastError());
  if (nslots > 0) {
    ocb_launch_finalize_partials(pl->partial, pl->nblocks, nslots, integrals, st);
    OCB_CUDA(cudaGetLastError());
  }
  return OCB_OK;
}

static bool plan_ring_capable(const ocb_plan* pl) {
  return pl && pl->fused && !pl->strain && pl->generic_blocks == 0 && pl->nslots > 0 && ocb_fused_ring_capable(pl->fused);
}

extern "C" int ocb_plan_ring_capable(const ocb_plan* pl) { return plan_ring_capable(pl) ? 1 : 0; }

// One slab's step of a peer-memory ring: the class-sum sweep with its edge tile rows gated on the neighbours' pushes, then
// the per-block partials summed and all-reduced over the ranks' sync blocks (include/oceanostics_b200.h).
extern "C" int ocb_plan_execute_ring(ocb_plan* pl, double* integrals, const ocb_ring* ring, int64_t step, void* stream) {
  OCB_REQUIRE(pl && integrals && ring, "ocb_plan_execute_ring: null argument");
  OCB_REQUIRE(ring->nranks >= 1 && ring->nranks <= OCB_MAX_RANKS && ring->rank >= 0 && ring->rank < ring->nranks,
              "ocb_plan_execute_ring: rank %d of %d (at most %d ranks)", ring->rank, ring->nranks, OCB_MAX_RANKS);
  OCB_REQUIRE(step >= 1 && ring->timeout_seconds > 0, "ocb_plan_execute_ring: step must be >= 1 and the timeout positive");
  for (int r = 0; r < ring->nranks; ++r) OCB_REQUIRE(ring->sync[r], "ocb_plan_execute_ring: sync block of rank %d is not mapped", r);
  if (!plan_ring_capable(pl))
    OCB_FAIL(OCB_ERR_UNSUPPORTED, "ocb_plan_execute_ring takes integral-only plans of the class-sum budget kernel over the whole y "
                                  "extent of a y-periodic slab; run the exchange, ocb_plan_execute and an all-reduce instead");
  OcbDeviceGuard device_guard(pl->device);
  cudaStream_t st = (cudaStream_t)stream;
  const long long cycles = ocb_cycles(ring->timeout_seconds);
  int rc = ocb_fused_execute_gated(pl->fused, pl->partial, (long long*)ring->sync[ring->rank], (long long)step, cycles, st);
  if (rc != OCB_OK) return rc;
  return ocb_launch_finalize_ring(pl->partial, pl->nblocks, pl->nslots, integrals, ring, (long long)step, cycles, st);
}

extern "C" int ocb_plan_destroy(ocb_plan* pl) {
  if (!pl) return OCB_OK;
  OcbDeviceGuard device_guard(pl->device);
  if (pl->fused) ocb_fused_destroy(pl->fused);
  if (pl->strain) ocb_strain_destroy(pl->strain);
  if (pl->ztab) cudaFree(pl->ztab);
  if (pl->partial) cudaFree(pl->partial);
  delete pl;
  return OCB_OK;
}

extern "C" int ocb_plan_info(const ocb_plan* pl, int32_t* n_launches, int32_t* variant, double* algorithmic_bytes) {
  OCB_REQUIRE(pl, "ocb_plan_info: null plan");
  if (n_launches) *n_launches = pl->n_launches;
  if (variant) *variant = pl->variant;
  if (algorithmic_bytes) *algorithmic_bytes = pl->algorithmic_bytes;
  return OCB_OK;
}
