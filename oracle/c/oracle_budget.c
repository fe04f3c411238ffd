/*
 * oracle_budget.c -- C/OpenMP twin of the NumPy oracle for the KE + tracer-variance budget (TEST
 * INFRASTRUCTURE / CPU BASELINE ONLY: used by tests/ and by bench.py's cpu_baseline and --impl reference legs,
 * never by the product library).
 *
 * It restates, per cell and WITHOUT any sharing of intermediates, the reference's kernel functions on a
 * regularly spaced RectilinearGrid (Periodic, Periodic, Bounded) with a ScalarDiffusivity closure:
 *   kinetic_energy_ccc                    src/KineticEnergyEquation.jl:37-39
 *   uᵢ∂ⱼuⱼuᵢᶜᶜᶜ  (Centered 2nd order)      src/KineticEnergyEquation.jl:147-152
 *   uᵢ∂ᵢpᶜᶜᶜ                              src/KineticEnergyEquation.jl:304-309
 *   uᵢbᵢᶜᶜᶜ                               src/KineticEnergyEquation.jl:362-367
 *   viscous_dissipation_rate_ccc          src/KineticEnergyEquation.jl:427-453
 *   tracer_variance_dissipation_rate_ccc  src/TracerVarianceEquation.jl:146-161
 *   c∇_dot_qᶜ                             src/TracerVarianceEquation.jl:93-98
 * with the Oceananigans operators (δ, ℑ, ∂, A, V, viscous / diffusive fluxes, centred momentum fluxes) restated
 * from the published Oceananigans v0.10x source, exactly as oracle/operators.py and oracle/kernels.py do.
 * Like the reference, every diagnostic is its own sweep ("one compute! per diagnostic") and every
 * interpolated edge quantity is re-evaluated at its 4 corners.  Threads: OpenMP over k-planes (the reference's
 * CPU architecture runs KernelAbstractions workgroups on Julia threads).
 */
#include <stddef.h>
#include <omp.h>

typedef struct {
  int Nx, Ny, Nz, H;            /* interior size and halo; parent arrays are (Nx+2H) x (Ny+2H) x (Nz(+1)+2H), i fastest */
  double dx, dy, dz;
  double nu, kappa;
  double ghat[3];
  const double *u, *v, *w, *b, *p;   /* parent arrays; w has Nz+1 interior levels */
} ob_state;

#define SX(S) ((ptrdiff_t)1)
#define SY(S) ((ptrdiff_t)((S)->Nx + 2 * (S)->H))
#define SZ(S) ((ptrdiff_t)((S)->Nx + 2 * (S)->H) * ((S)->Ny + 2 * (S)->H))
/* 1-based halo-aware access */
#define AT(S, f, i, j, k) ((S)->f[((i) + (S)->H - 1) + ((j) + (S)->H - 1) * SY(S) + ((k) + (S)->H - 1) * SZ(S)])

typedef double (*cellfn)(const ob_state*, int, int, int);

/* ---- generic operators on per-point functions (Oceananigans Operators) ------------------------------------ */
static inline double dx_c(cellfn f, const ob_state* S, int i, int j, int k) { return f(S, i + 1, j, k) - f(S, i, j, k); }
static inline double dx_f(cellfn f, const ob_state* S, int i, int j, int k) { return f(S, i, j, k) - f(S, i - 1, j, k); }
static inline double dy_c(cellfn f, const ob_state* S, int i, int j, int k) { return f(S, i, j + 1, k) - f(S, i, j, k); }
static inline double dy_f(cellfn f, const ob_state* S, int i, int j, int k) { return f(S, i, j, k) - f(S, i, j - 1, k); }
static inline double dz_c(cellfn f, const ob_state* S, int i, int j, int k) { return f(S, i, j, k + 1) - f(S, i, j, k); }
static inline double dz_f(cellfn f, const ob_state* S, int i, int j, int k) { return f(S, i, j, k) - f(S, i, j, k - 1); }
static inline double ix_c(cellfn f, const ob_state* S, int i, int j, int k) { return (f(S, i, j, k) + f(S, i + 1, j, k)) / 2; }
static inline double ix_f(cellfn f, const ob_state* S, int i, int j, int k) { return (f(S, i - 1, j, k) + f(S, i, j, k)) / 2; }
static inline double iy_c(cellfn f, const ob_state* S, int i, int j, int k) { return (f(S, i, j, k) + f(S, i, j + 1, k)) / 2; }
static inline double iy_f(cellfn f, const ob_state* S, int i, int j, int k) { return (f(S, i, j - 1, k) + f(S, i, j, k)) / 2; }
static inline double iz_c(cellfn f, const ob_state* S, int i, int j, int k) { return (f(S, i, j, k) + f(S, i, j, k + 1)) / 2; }
static inline double iz_f(cellfn f, const ob_state* S, int i, int j, int k) { return (f(S, i, j, k - 1) + f(S, i, j, k)) / 2; }
/* double interpolations: the later direction is the outer operator */
static inline double ixy_cc(cellfn f, const ob_state* S, int i, int j, int k) { return (ix_c(f, S, i, j, k) + ix_c(f, S, i, j + 1, k)) / 2; }
static inline double ixz_cc(cellfn f, const ob_state* S, int i, int j, int k) { return (ix_c(f, S, i, j, k) + ix_c(f, S, i, j, k + 1)) / 2; }
static inline double iyz_cc(cellfn f, const ob_state* S, int i, int j, int k) { return (iy_c(f, S, i, j, k) + iy_c(f, S, i, j, k + 1)) / 2; }

/* fields as per-point functions */
static double fu(const ob_state* S, int i, int j, int k) { return AT(S, u, i, j, k); }
static double fv(const ob_state* S, int i, int j, int k) { return AT(S, v, i, j, k); }
static double fw(const ob_state* S, int i, int j, int k) { return AT(S, w, i, j, k); }
static double fb(const ob_state* S, int i, int j, int k) { return AT(S, b, i, j, k); }
static double fp(const ob_state* S, int i, int j, int k) { return AT(S, p, i, j, k); }
static double fu2(const ob_state* S, int i, int j, int k) { double a = AT(S, u, i, j, k); return a * a; }
static double fv2(const ob_state* S, int i, int j, int k) { double a = AT(S, v, i, j, k); return a * a; }
static double fw2(const ob_state* S, int i, int j, int k) { double a = AT(S, w, i, j, k); return a * a; }

/* metrics (regular spacing) */
#define AX(S) ((S)->dy * (S)->dz)
#define AY(S) ((S)->dx * (S)->dz)
#define AZ(S) ((S)->dx * (S)->dy)
#define VOL(S) ((S)->dx * (S)->dy * (S)->dz)

/* ---- kinetic_energy_ccc ------------------------------------------------------------------------------------- */
static double k_ke(const ob_state* S, int i, int j, int k) {
  return (ix_c(fu2, S, i, j, k) + iy_c(fv2, S, i, j, k) + iz_c(fw2, S, i, j, k)) / 2;
}

/* ---- strain rates and viscous fluxes (TurbulenceClosures, isotropic ScalarDiffusivity) ---------------------- */
static double S11(const ob_state* S, int i, int j, int k) { return dx_c(fu, S, i, j, k) / S->dx; }
static double S22(const ob_state* S, int i, int j, int k) { return dy_c(fv, S, i, j, k) / S->dy; }
static double S33(const ob_state* S, int i, int j, int k) { return dz_c(fw, S, i, j, k) / S->dz; }
static double S12(const ob_state* S, int i, int j, int k) { return (dy_f(fu, S, i, j, k) / S->dy + dx_f(fv, S, i, j, k) / S->dx) / 2; }
static double S13(const ob_state* S, int i, int j, int k) { return (dz_f(fu, S, i, j, k) / S->dz + dx_f(fw, S, i, j, k) / S->dx) / 2; }
static double S23(const ob_state* S, int i, int j, int k) { return (dz_f(fv, S, i, j, k) / S->dz + dy_f(fw, S, i, j, k) / S->dy) / 2; }
static double A_du_t11(const ob_state* S, int i, int j, int k) { return -AX(S) * dx_c(fu, S, i, j, k) * (-2 * (S->nu * S11(S, i, j, k))); }
static double A_du_t12(const ob_state* S, int i, int j, int k) { return -AY(S) * dy_f(fu, S, i, j, k) * (-2 * (S->nu * S12(S, i, j, k))); }
static double A_du_t13(const ob_state* S, int i, int j, int k) { return -AZ(S) * dz_f(fu, S, i, j, k) * (-2 * (S->nu * S13(S, i, j, k))); }
static double A_dv_t21(const ob_state* S, int i, int j, int k) { return -AX(S) * dx_f(fv, S, i, j, k) * (-2 * (S->nu * S12(S, i, j, k))); }
static double A_dv_t22(const ob_state* S, int i, int j, int k) { return -AY(S) * dy_c(fv, S, i, j, k) * (-2 * (S->nu * S22(S, i, j, k))); }
static double A_dv_t23(const ob_state* S, int i, int j, int k) { return -AZ(S) * dz_f(fv, S, i, j, k) * (-2 * (S->nu * S23(S, i, j, k))); }
static double A_dw_t31(const ob_state* S, int i, int j, int k) { return -AX(S) * dx_f(fw, S, i, j, k) * (-2 * (S->nu * S13(S, i, j, k))); }
static double A_dw_t32(const ob_state* S, int i, int j, int k) { return -AY(S) * dy_f(fw, S, i, j, k) * (-2 * (S->nu * S23(S, i, j, k))); }
static double A_dw_t33(const ob_state* S, int i, int j, int k) { return -AZ(S) * dz_c(fw, S, i, j, k) * (-2 * (S->nu * S33(S, i, j, k))); }
/* viscous_dissipation_rate_ccc: nine terms, six of them interpolated from their edges (4 corner evaluations each) */
static double k_eps(const ob_state* S, int i, int j, int k) {
  return (A_du_t11(S, i, j, k) + ixy_cc(A_du_t12, S, i, j, k) + ixz_cc(A_du_t13, S, i, j, k) +
          ixy_cc(A_dv_t21, S, i, j, k) + A_dv_t22(S, i, j, k) + iyz_cc(A_dv_t23, S, i, j, k) +
          ixz_cc(A_dw_t31, S, i, j, k) + iyz_cc(A_dw_t32, S, i, j, k) + A_dw_t33(S, i, j, k)) / VOL(S);
}

/* ---- pressure redistribution and buoyancy production ---------------------------------------------------------- */
static double u_dxp(const ob_state* S, int i, int j, int k) { return AT(S, u, i, j, k) * (dx_f(fp, S, i, j, k) / S->dx); }
static double v_dyp(const ob_state* S, int i, int j, int k) { return AT(S, v, i, j, k) * (dy_f(fp, S, i, j, k) / S->dy); }
static double w_dzp(const ob_state* S, int i, int j, int k) { return AT(S, w, i, j, k) * (dz_f(fp, S, i, j, k) / S->dz); }
static double k_pr(const ob_state* S, int i, int j, int k) { return ix_c(u_dxp, S, i, j, k) + iy_c(v_dyp, S, i, j, k) + iz_c(w_dzp, S, i, j, k); }
static double u_bx(const ob_state* S, int i, int j, int k) { return AT(S, u, i, j, k) * (S->ghat[0] * ix_f(fb, S, i, j, k)); }
static double v_by(const ob_state* S, int i, int j, int k) { return AT(S, v, i, j, k) * (S->ghat[1] * iy_f(fb, S, i, j, k)); }
static double w_bz(const ob_state* S, int i, int j, int k) { return AT(S, w, i, j, k) * (S->ghat[2] * iz_f(fb, S, i, j, k)); }
static double k_wb(const ob_state* S, int i, int j, int k) { return ix_c(u_bx, S, i, j, k) + iy_c(v_by, S, i, j, k) + iz_c(w_bz, S, i, j, k); }

/* ---- Centered second-order momentum advection (Advection/momentum_advection_operators.jl) -------------------- */
static double Axu(const ob_state* S, int i, int j, int k) { return AX(S) * AT(S, u, i, j, k); }
static double Ayv(const ob_state* S, int i, int j, int k) { return AY(S) * AT(S, v, i, j, k); }
static double Azw(const ob_state* S, int i, int j, int k) { return AZ(S) * AT(S, w, i, j, k); }
static double flux_Uu(const ob_state* S, int i, int j, int k) { return ix_c(Axu, S, i, j, k) * ix_c(fu, S, i, j, k); }
static double flux_Vu(const ob_state* S, int i, int j, int k) { return ix_f(Ayv, S, i, j, k) * iy_f(fu, S, i, j, k); }
static double flux_Wu(const ob_state* S, int i, int j, int k) { return ix_f(Azw, S, i, j, k) * iz_f(fu, S, i, j, k); }
static double flux_Uv(const ob_state* S, int i, int j, int k) { return iy_f(Axu, S, i, j, k) * ix_f(fv, S, i, j, k); }
static double flux_Vv(const ob_state* S, int i, int j, int k) { return iy_c(Ayv, S, i, j, k) * iy_c(fv, S, i, j, k); }
static double flux_Wv(const ob_state* S, int i, int j, int k) { return iy_f(Azw, S, i, j, k) * iz_f(fv, S, i, j, k); }
static double flux_Uw(const ob_state* S, int i, int j, int k) { return iz_f(Axu, S, i, j, k) * ix_f(fw, S, i, j, k); }
static double flux_Vw(const ob_state* S, int i, int j, int k) { return iz_f(Ayv, S, i, j, k) * iy_f(fw, S, i, j, k); }
static double flux_Ww(const ob_state* S, int i, int j, int k) { return iz_c(Azw, S, i, j, k) * iz_c(fw, S, i, j, k); }
static double u_divUu(const ob_state* S, int i, int j, int k) {
  return AT(S, u, i, j, k) * (1 / VOL(S) * (dx_f(flux_Uu, S, i, j, k) + dy_c(flux_Vu, S, i, j, k) + dz_c(flux_Wu, S, i, j, k)));
}
static double v_divUv(const ob_state* S, int i, int j, int k) {
  return AT(S, v, i, j, k) * (1 / VOL(S) * (dx_c(flux_Uv, S, i, j, k) + dy_f(flux_Vv, S, i, j, k) + dz_c(flux_Wv, S, i, j, k)));
}
static double w_divUw(const ob_state* S, int i, int j, int k) {
  return AT(S, w, i, j, k) * (1 / VOL(S) * (dx_c(flux_Uw, S, i, j, k) + dy_c(flux_Vw, S, i, j, k) + dz_f(flux_Ww, S, i, j, k)));
}
static double k_adv(const ob_state* S, int i, int j, int k) { return ix_c(u_divUu, S, i, j, k) + iy_c(v_divUv, S, i, j, k) + iz_c(w_divUw, S, i, j, k); }

/* ---- tracer variance --------------------------------------------------------------------------------------------- */
static double q1(const ob_state* S, int i, int j, int k) { return -(S->kappa * (dx_f(fb, S, i, j, k) / S->dx)); }
static double q2(const ob_state* S, int i, int j, int k) { return -(S->kappa * (dy_f(fb, S, i, j, k) / S->dy)); }
static double q3(const ob_state* S, int i, int j, int k) { return -(S->kappa * (dz_f(fb, S, i, j, k) / S->dz)); }
static double Ax_dc_q1(const ob_state* S, int i, int j, int k) { return -AX(S) * dx_f(fb, S, i, j, k) * q1(S, i, j, k); }
static double Ay_dc_q2(const ob_state* S, int i, int j, int k) { return -AY(S) * dy_f(fb, S, i, j, k) * q2(S, i, j, k); }
static double Az_dc_q3(const ob_state* S, int i, int j, int k) { return -AZ(S) * dz_f(fb, S, i, j, k) * q3(S, i, j, k); }
static double k_chi(const ob_state* S, int i, int j, int k) {
  return 2 * (ix_c(Ax_dc_q1, S, i, j, k) + iy_c(Ay_dc_q2, S, i, j, k) + iz_c(Az_dc_q3, S, i, j, k)) / VOL(S);
}
static double Ax_q1(const ob_state* S, int i, int j, int k) { return AX(S) * q1(S, i, j, k); }
static double Ay_q2(const ob_state* S, int i, int j, int k) { return AY(S) * q2(S, i, j, k); }
static double Az_q3(const ob_state* S, int i, int j, int k) { return AZ(S) * q3(S, i, j, k); }
static double k_cdiff(const ob_state* S, int i, int j, int k) {
  return 2 * AT(S, b, i, j, k) * (1 / VOL(S) * (dx_c(Ax_q1, S, i, j, k) + dy_c(Ay_q2, S, i, j, k) + dz_c(Az_q3, S, i, j, k)));
}

static cellfn pick(int diag) {
  switch (diag) {
    case 0: return k_ke;
    case 1: return k_adv;
    case 2: return k_pr;
    case 3: return k_wb;
    case 4: return k_eps;
    case 5: return k_chi;
    case 6: return k_cdiff;
    default: return 0;
  }
}

/* Field(op): interior values, column-major (i fastest), one sweep */
int ob_field(const ob_state* S, int diag, double* out) {
  if (!pick(diag)) return 1;
#pragma omp parallel for schedule(static)
  for (int k = 1; k <= S->Nz; ++k)
    for (int j = 1; j <= S->Ny; ++j)
      for (int i = 1; i <= S->Nx; ++i) {
        double v;
        switch (diag) {   /* direct calls so the compiler can inline each kernel function as Julia does */
          case 0: v = k_ke(S, i, j, k); break;
          case 1: v = k_adv(S, i, j, k); break;
          case 2: v = k_pr(S, i, j, k); break;
          case 3: v = k_wb(S, i, j, k); break;
          case 4: v = k_eps(S, i, j, k); break;
          case 5: v = k_chi(S, i, j, k); break;
          default: v = k_cdiff(S, i, j, k); break;
        }
        out[(size_t)(i - 1) + (size_t)(j - 1) * S->Nx + (size_t)(k - 1) * S->Nx * S->Ny] = v;
      }
  return 0;
}

/* sum(Integral(op)) = Σ op·V : the operand is not materialised (Reduction(sum!, op * V)), one sweep per integral */
double ob_integral(const ob_state* S, int diag) {
  double total = 0.0;
  const double V = VOL(S);
#pragma omp parallel for schedule(static) reduction(+ : total)
  for (int k = 1; k <= S->Nz; ++k) {
    double plane = 0.0;
    for (int j = 1; j <= S->Ny; ++j)
      for (int i = 1; i <= S->Nx; ++i) {
        double v;
        switch (diag) {
          case 0: v = k_ke(S, i, j, k); break;
          case 1: v = k_adv(S, i, j, k); break;
          case 2: v = k_pr(S, i, j, k); break;
          case 3: v = k_wb(S, i, j, k); break;
          case 4: v = k_eps(S, i, j, k); break;
          case 5: v = k_chi(S, i, j, k); break;
          default: v = k_cdiff(S, i, j, k); break;
        }
        plane += v * V;
      }
    total += plane;
  }
  return total;
}

int ob_num_threads(void) { return omp_get_max_threads(); }
/* torchrun exports OMP_NUM_THREADS=1 to its workers: the CPU arm of the bench asks for the host's cores explicitly */
void ob_set_num_threads(int n) { if (n > 0) omp_set_num_threads(n); }
