// TEST HARNESS ONLY (never linked into liboceanostics_b200.so): runs the library's own per-cell
// evaluators (csrc/ocb_diag.cuh, the exact templates the CUDA sweep kernel instantiates) on the HOST over
// host arrays, so the formulas can be checked against the NumPy oracle in the GPU-less build container.
// The product path has no CPU route; this file exists so a wrong stencil is caught before GPU time is spent.
#include "../../oceanostics.jl_b200/csrc/ocb_sweep_plan.cuh"
#include <vector>

static thread_local char g_err[1024] = "";
void ocb_set_error(const char* fmt, ...) {
  va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap);
}
extern "C" const char* hc_last_error(void) { return g_err; }

template <class T>
static int run(const ocb_grid* grid, const ocb_sweep_inputs* in, const ocb_request* reqs, int n, double* integrals) {
  const int Hz = grid->H[2], nt = grid->N[2] + 2 * Hz + 1;
  std::vector<T> tab(6 * (size_t)nt);
  for (int m = 0; m < nt; ++m) {
    const int k = m + 1 - Hz;
    T dzc, dzf, zc, zf;
    if (grid->dzc) {
      const int mc = m < nt - 1 ? m : nt - 2;
      dzc = ((const T*)grid->dzc)[mc]; zc = grid->zc ? ((const T*)grid->zc)[mc] : T(0);
      dzf = ((const T*)grid->dzf)[m];  zf = grid->zf ? ((const T*)grid->zf)[m] : T(0);
    } else {
      dzc = dzf = T(grid->d[2]); zf = T(grid->z_bottom) + dzc * T(k - 1); zc = T(grid->z_bottom) + dzc * (T(k) - T(0.5));
    }
    tab[0 * nt + m] = dzc; tab[1 * nt + m] = dzf; tab[2 * nt + m] = T(1) / dzc; tab[3 * nt + m] = T(1) / dzf;
    tab[4 * nt + m] = zc; tab[5 * nt + m] = zf;
  }
  ZTables<T> zt;
  const int off = Hz - 1;
  zt.dzc = tab.data() + off; zt.dzf = tab.data() + nt + off; zt.rdzc = tab.data() + 2 * nt + off;
  zt.rdzf = tab.data() + 3 * nt + off; zt.zc = tab.data() + 4 * nt + off; zt.zf = tab.data() + 5 * nt + off;
  SweepParams<T> P;
  int rc = ocb_build_sweep_params<T>(grid, in, reqs, n, zt, &P);
  if (rc != OCB_OK) return rc;
  for (int r = 0; r < P.nreq; ++r) {
    const DReq& R = P.req[r];
    const bool pert = (R.flags & OCB_REQ_PERTURBATION) != 0;
    double acc = 0.0;
    // same cell-relative addressing as the CUDA kernel: shift the operand views to the cell, evaluate at the origin
    for (int j = R.lo[1]; j <= R.hi[1]; ++j)
      for (int i = R.lo[0]; i <= R.hi[0]; ++i) {
        SweepIn<T> L = P.in;
        DGrid<T> G = P.g;
        ocb_shift_inputs(L, G, i, j, R.lo[2]);
        const Ctx<T> X(G, L, pert);
        for (int k = R.lo[2]; k <= R.hi[2]; ++k) {
          T val = T(0);
          switch (R.diag) {
#define HC_CASE(D) case D: val = Eval<T, D>::at(X, R.flags, 0, 0, 0); break;
            OCB_FOR_EACH_DIAG(HC_CASE)
#undef HC_CASE
          }
          if (R.out) ((T*)R.out)[i * R.os[0] + j * R.os[1] + k * R.os[2]] = val;
          if (R.slot >= 0) acc += (double)val * ((double)P.g.dx * (double)P.g.dy * (double)(R.loc[2] == OCB_FACE ? G.dzf[0] : G.dzc[0]));
          ocb_shift_inputs(L, G, 0, 0, 1);
        }
      }
    if (R.slot >= 0 && integrals) integrals[R.slot] = acc;
  }
  return OCB_OK;
}

extern "C" int hc_sweep(const ocb_grid* grid, const ocb_sweep_inputs* in, const ocb_request* reqs, int n, double* integrals) {
  if (!grid) return OCB_ERR_INVALID;
  return grid->dtype == OCB_F64 ? run<double>(grid, in, reqs, n, integrals) : run<float>(grid, in, reqs, n, integrals);
}
