// Per-cell evaluation of every diagnostic of the fused stencil sweep (K1), as host+device templates.
//
// These are the B200 library's own formulations of the kernel functions the reference evaluates
// through Oceananigans' operator algebra (file:line of each reference function is given at its
// evaluator).  They are written on explicit index offsets with reciprocal metrics, each edge / face
// quantity formed once per evaluation -- mathematically equal to the reference, not bit-equal
// (the contract is rtol 1e-12 in Float64, BASELINE.json north_star).
//
// The same templates are compiled (a) into the generic CUDA sweep kernel (ocb_sweep.cu) and (b), for
// testing only, into tests/hostcheck/ where they run on the host against the NumPy oracle.
#pragma once
#include "ocb_common.cuh"

#define OCB_MAX_REQ 16

template <class T>
struct SweepIn {
  HFld<T> u, v, w, b, p, U, V, W, B;
  HFld<T> aux[OCB_N_AUX];
  T nu_c; int n_nu; HFld<T> nu_f[OCB_MAX_CLOSURE_FIELDS];
  T ka_c; int n_ka; HFld<T> ka_f[OCB_MAX_CLOSURE_FIELDS]; T ka_s[OCB_MAX_CLOSURE_FIELDS];
  T f[3], dir[3], f_dir, ghat[3], ro_bg[6], les[4];
};

struct DReq {
  int diag, flags;
  void* out;             // pre-offset: index (i,j,k) lives at out[i*os[0] + j*os[1] + k*os[2]]
  long long os[3];
  int lo[3], hi[3];      // inclusive 1-based index window to evaluate
  int slot;
  int loc[3];
};

template <class T>
struct SweepParams {
  DGrid<T> g;
  SweepIn<T> in;
  int nreq;
  DReq req[OCB_MAX_REQ];
  double* partial;       // [nblocks][nslots]
  int nslots;
  int lo[3], hi[3];      // union of the request windows
  int kchunk;
};

// ---- evaluation context: the operands one request sees ----------------------------------------------
template <class T>
struct Ctx {
  const DGrid<T>& g;
  const SweepIn<T>& in;
  bool pert;
  OCB_X Ctx(const DGrid<T>& g_, const SweepIn<T>& in_, bool p_) : g(g_), in(in_), pert(p_) {}
  // velocities / tracer as this request sees them (u - U when the perturbation flag is set)
  OCB_X T u(int i, int j, int k) const { T a = in.u(i, j, k); return pert ? a - in.U(i, j, k) : a; }
  OCB_X T v(int i, int j, int k) const { T a = in.v(i, j, k); return pert ? a - in.V(i, j, k) : a; }
  OCB_X T w(int i, int j, int k) const { T a = in.w(i, j, k); return pert ? a - in.W(i, j, k) : a; }
  OCB_X T c(int i, int j, int k) const { T a = in.b(i, j, k); return pert ? a - in.B(i, j, k) : a; }
  // z metrics
  OCB_X T dzc(int k) const { return OCB_LD(g.dzc + k); }
  OCB_X T dzf(int k) const { return OCB_LD(g.dzf + k); }
  OCB_X T rdzc(int k) const { return OCB_LD(g.rdzc + k); }
  OCB_X T rdzf(int k) const { return OCB_LD(g.rdzf + k); }
  // closure viscosity at the four stress locations: the constant plus the interpolated eddy fields
  OCB_X T nu_ccc(int i, int j, int k) const {
    T s = in.nu_c;
    _Pragma("unroll") for (int f = 0; f < OCB_MAX_CLOSURE_FIELDS; ++f) if (f < in.n_nu) s += in.nu_f[f](i, j, k);
    return s;
  }
  OCB_X T nu_ffc(int i, int j, int k) const {
    T s = in.nu_c;
    _Pragma("unroll") for (int f = 0; f < OCB_MAX_CLOSURE_FIELDS; ++f) {
      if (f >= in.n_nu) continue;
      const HFld<T>& n = in.nu_f[f];
      s += T(0.25) * ((n(i - 1, j - 1, k) + n(i, j - 1, k)) + (n(i - 1, j, k) + n(i, j, k)));
    }
    return s;
  }
  OCB_X T nu_fcf(int i, int j, int k) const {
    T s = in.nu_c;
    _Pragma("unroll") for (int f = 0; f < OCB_MAX_CLOSURE_FIELDS; ++f) {
      if (f >= in.n_nu) continue;
      const HFld<T>& n = in.nu_f[f];
      s += T(0.25) * ((n(i - 1, j, k - 1) + n(i, j, k - 1)) + (n(i - 1, j, k) + n(i, j, k)));
    }
    return s;
  }
  OCB_X T nu_cff(int i, int j, int k) const {
    T s = in.nu_c;
    _Pragma("unroll") for (int f = 0; f < OCB_MAX_CLOSURE_FIELDS; ++f) {
      if (f >= in.n_nu) continue;
      const HFld<T>& n = in.nu_f[f];
      s += T(0.25) * ((n(i, j - 1, k - 1) + n(i, j, k - 1)) + (n(i, j - 1, k) + n(i, j, k)));
    }
    return s;
  }
  // tracer diffusivity at the three flux faces
  OCB_X T ka_fcc(int i, int j, int k) const {
    T s = in.ka_c;
    _Pragma("unroll") for (int f = 0; f < OCB_MAX_CLOSURE_FIELDS; ++f) if (f < in.n_ka) s += in.ka_s[f] * (T(0.5) * (in.ka_f[f](i - 1, j, k) + in.ka_f[f](i, j, k)));
    return s;
  }
  OCB_X T ka_cfc(int i, int j, int k) const {
    T s = in.ka_c;
    _Pragma("unroll") for (int f = 0; f < OCB_MAX_CLOSURE_FIELDS; ++f) if (f < in.n_ka) s += in.ka_s[f] * (T(0.5) * (in.ka_f[f](i, j - 1, k) + in.ka_f[f](i, j, k)));
    return s;
  }
  OCB_X T ka_ccf(int i, int j, int k) const {
    T s = in.ka_c;
    _Pragma("unroll") for (int f = 0; f < OCB_MAX_CLOSURE_FIELDS; ++f) if (f < in.n_ka) s += in.ka_s[f] * (T(0.5) * (in.ka_f[f](i, j, k - 1) + in.ka_f[f](i, j, k)));
    return s;
  }

  // ---- velocity-gradient building blocks ---------------------------------------------------------------
  // diagonal strain rates at ccc
  OCB_X T s11(int i, int j, int k) const { return (u(i + 1, j, k) - u(i, j, k)) * g.rdx; }
  OCB_X T s22(int i, int j, int k) const { return (v(i, j + 1, k) - v(i, j, k)) * g.rdy; }
  OCB_X T s33(int i, int j, int k) const { return (w(i, j, k + 1) - w(i, j, k)) * rdzc(k); }
  // the six cross derivatives at their edges
  OCB_X T dyu_ffc(int i, int j, int k) const { return (u(i, j, k) - u(i, j - 1, k)) * g.rdy; }
  OCB_X T dxv_ffc(int i, int j, int k) const { return (v(i, j, k) - v(i - 1, j, k)) * g.rdx; }
  OCB_X T dzu_fcf(int i, int j, int k) const { return (u(i, j, k) - u(i, j, k - 1)) * rdzf(k); }
  OCB_X T dxw_fcf(int i, int j, int k) const { return (w(i, j, k) - w(i - 1, j, k)) * g.rdx; }
  OCB_X T dzv_cff(int i, int j, int k) const { return (v(i, j, k) - v(i, j, k - 1)) * rdzf(k); }
  OCB_X T dyw_cff(int i, int j, int k) const { return (w(i, j, k) - w(i, j - 1, k)) * g.rdy; }
  OCB_X T s12(int i, int j, int k) const { return T(0.5) * (dyu_ffc(i, j, k) + dxv_ffc(i, j, k)); }
  OCB_X T s13(int i, int j, int k) const { return T(0.5) * (dzu_fcf(i, j, k) + dxw_fcf(i, j, k)); }
  OCB_X T s23(int i, int j, int k) const { return T(0.5) * (dzv_cff(i, j, k) + dyw_cff(i, j, k)); }
};

// four-point means from an edge location to the cell centre
#define OCB_MEAN_XY(f, i, j, k) (T(0.25) * (((f)((i), (j), (k)) + (f)((i) + 1, (j), (k))) + ((f)((i), (j) + 1, (k)) + (f)((i) + 1, (j) + 1, (k)))))
#define OCB_MEAN_XZ(f, i, j, k) (T(0.25) * (((f)((i), (j), (k)) + (f)((i) + 1, (j), (k))) + ((f)((i), (j), (k) + 1) + (f)((i) + 1, (j), (k) + 1))))
#define OCB_MEAN_YZ(f, i, j, k) (T(0.25) * (((f)((i), (j), (k)) + (f)((i), (j) + 1, (k))) + ((f)((i), (j), (k) + 1) + (f)((i), (j) + 1, (k) + 1))))

template <class T> OCB_X T ocb_sq(T x) { return x * x; }

// =====================================================================================================
// evaluators: one per OCB_DIAG_* id
// =====================================================================================================
template <class T, int DIAG>
struct Eval;   // Eval<T, DIAG>::at(ctx, flags, i, j, k)

#define OCB_EVAL(DIAG) template <class T> struct Eval<T, DIAG> { static OCB_X T at(const Ctx<T>& X, int flags, int i, int j, int k)

// kinetic_energy_ccc  (src/KineticEnergyEquation.jl:37-39): half the sum of the face-averaged squares
template <class T> OCB_X T ocb_ke(const Ctx<T>& X, int i, int j, int k) {
  T a = ocb_sq(X.u(i, j, k)) + ocb_sq(X.u(i + 1, j, k));
  T b = ocb_sq(X.v(i, j, k)) + ocb_sq(X.v(i, j + 1, k));
  T c = ocb_sq(X.w(i, j, k)) + ocb_sq(X.w(i, j, k + 1));
  return T(0.25) * (a + b + c);
}
OCB_EVAL(OCB_DIAG_KE) { return ocb_ke(X, i, j, k); } };

// turbulent_kinetic_energy_ccc  (src/TurbulentKineticEnergyEquation.jl:24-30): always about (U, V, W)
OCB_EVAL(OCB_DIAG_TKE) {
  Ctx<T> P(X.g, X.in, true);
  return ocb_ke(P, i, j, k);
} };

// ---- uᵢ∂ⱼuⱼuᵢᶜᶜᶜ, Centered second order  (src/KineticEnergyEquation.jl:147-152) -----------------------
// momentum fluxes: (area-weighted advecting velocity averaged to the flux point) x (advected velocity there)
// Centered(order = N) symmetric interpolation [OCN-mem Advection/centered_reconstruction.jl]: the two-point mean (order 2) or
// (-1, 7, 7, -1)/12 over the four nearest points (order 4; uniform-grid coefficients on every axis, as `Centered(order = 4)` built
// without a grid has them); along a Bounded axis order 4 falls back to order 2 where the stencil would reach across a wall
// (outside_symmetric_halo: face index m keeps it for 3 <= m <= N-1, centre index for 2 <= m <= N-1).
template <class T> struct AdvFlux {
  const Ctx<T>& X;
  bool o4;
  OCB_X AdvFlux(const Ctx<T>& x, int flags = 0) : X(x), o4(flags & OCB_REQ_CENTERED4) {}
  OCB_X T axu(int i, int j, int k) const { return X.g.dy * X.dzc(k) * X.u(i, j, k); }   // Ax_fcc u
  OCB_X T ayv(int i, int j, int k) const { return X.g.dx * X.dzc(k) * X.v(i, j, k); }   // Ay_cfc v
  OCB_X T azw(int i, int j, int k) const { return X.g.dx * X.g.dy * X.w(i, j, k); }     // Az_ccf w
  // does the order-4 stencil along axis d fit at index m (absolute = origin + m)?  to_face: interpolation to a face index
  OCB_X bool wide(int d, int m, bool to_face) const {
    if (!o4) return false;
    const int topo = d == 0 ? X.g.tx : (d == 1 ? X.g.ty : X.g.tz);
    if (topo != OCB_BOUNDED) return true;
    const int N = d == 0 ? X.g.Nx : (d == 1 ? X.g.Ny : X.g.Nz);
    const int a = m + (d == 0 ? X.g.oi : (d == 1 ? X.g.oj : X.g.ok));
    return a >= (to_face ? 3 : 2) && a <= N - 1;
  }
  // interpolation of four samples (s_-1, s_0, s_1, s_2 around the target) -- the middle two for order 2
  OCB_X T mix(bool w4, T sm, T s0, T s1, T s2) const { return w4 ? (T(7) * (s0 + s1) - (sm + s2)) * T(1.0 / 12.0) : T(0.5) * (s0 + s1); }
#define OCB_I4(FN, AXIS, TOFACE, IDX, LO, ARG)                                                             \
  OCB_X T FN(int i, int j, int k) const {                                                                   \
    const bool w4 = wide(AXIS, IDX, TOFACE);                                                                \
    return mix(w4, w4 ? ARG(LO - 1) : T(0), ARG(LO), ARG(LO + 1), w4 ? ARG(LO + 2) : T(0));                \
  }
  // advecting (area-weighted) velocities and advected velocities at the nine flux points; to a centre: samples m, m+1 (LO = 0),
  // to a face: samples m-1, m (LO = -1)
#define A_(s) axu(i + (s), j, k)
  OCB_I4(Ucx, 0, false, i, 0, A_)     // Ax u -> ccc along x
#undef A_
#define A_(s) axu(i, j + (s), k)
  OCB_I4(Ufy, 1, true, j, -1, A_)     // Ax u -> ffc along y
#undef A_
#define A_(s) axu(i, j, k + (s))
  OCB_I4(Ufz, 2, true, k, -1, A_)     // Ax u -> fcf along z
#undef A_
#define A_(s) ayv(i + (s), j, k)
  OCB_I4(Vfx, 0, true, i, -1, A_)     // Ay v -> ffc along x
#undef A_
#define A_(s) ayv(i, j + (s), k)
  OCB_I4(Vcy, 1, false, j, 0, A_)     // Ay v -> ccc along y
#undef A_
#define A_(s) ayv(i, j, k + (s))
  OCB_I4(Vfz, 2, true, k, -1, A_)     // Ay v -> cff along z
#undef A_
#define A_(s) azw(i + (s), j, k)
  OCB_I4(Wfx, 0, true, i, -1, A_)     // Az w -> fcf along x
#undef A_
#define A_(s) azw(i, j + (s), k)
  OCB_I4(Wfy, 1, true, j, -1, A_)     // Az w -> cff along y
#undef A_
#define A_(s) azw(i, j, k + (s))
  OCB_I4(Wcz, 2, false, k, 0, A_)     // Az w -> ccc along z
#undef A_
#define A_(s) X.u(i + (s), j, k)
  OCB_I4(ucx, 0, false, i, 0, A_)
#undef A_
#define A_(s) X.u(i, j + (s), k)
  OCB_I4(ufy, 1, true, j, -1, A_)
#undef A_
#define A_(s) X.u(i, j, k + (s))
  OCB_I4(ufz, 2, true, k, -1, A_)
#undef A_
#define A_(s) X.v(i + (s), j, k)
  OCB_I4(vfx, 0, true, i, -1, A_)
#undef A_
#define A_(s) X.v(i, j + (s), k)
  OCB_I4(vcy, 1, false, j, 0, A_)
#undef A_
#define A_(s) X.v(i, j, k + (s))
  OCB_I4(vfz, 2, true, k, -1, A_)
#undef A_
#define A_(s) X.w(i + (s), j, k)
  OCB_I4(wfx, 0, true, i, -1, A_)
#undef A_
#define A_(s) X.w(i, j + (s), k)
  OCB_I4(wfy, 1, true, j, -1, A_)
#undef A_
#define A_(s) X.w(i, j, k + (s))
  OCB_I4(wcz, 2, false, k, 0, A_)
#undef A_
#undef OCB_I4
  // u-momentum fluxes
  OCB_X T Uu(int i, int j, int k) const { return Ucx(i, j, k) * ucx(i, j, k); }   // ccc
  OCB_X T Vu(int i, int j, int k) const { return Vfx(i, j, k) * ufy(i, j, k); }   // ffc
  OCB_X T Wu(int i, int j, int k) const { return Wfx(i, j, k) * ufz(i, j, k); }   // fcf
  // v-momentum
  OCB_X T Uv(int i, int j, int k) const { return Ufy(i, j, k) * vfx(i, j, k); }   // ffc
  OCB_X T Vv(int i, int j, int k) const { return Vcy(i, j, k) * vcy(i, j, k); }   // ccc
  OCB_X T Wv(int i, int j, int k) const { return Wfy(i, j, k) * vfz(i, j, k); }   // cff
  // w-momentum
  OCB_X T Uw(int i, int j, int k) const { return Ufz(i, j, k) * wfx(i, j, k); }   // fcf
  OCB_X T Vw(int i, int j, int k) const { return Vfz(i, j, k) * wfy(i, j, k); }   // cff
  OCB_X T Ww(int i, int j, int k) const { return Wcz(i, j, k) * wcz(i, j, k); }   // ccc
  // flux divergences at the velocity points
  OCB_X T divu(int i, int j, int k) const {   // fcc
    T s = (Uu(i, j, k) - Uu(i - 1, j, k)) + (Vu(i, j + 1, k) - Vu(i, j, k)) + (Wu(i, j, k + 1) - Wu(i, j, k));
    return s * (X.g.rdx * X.g.rdy * X.rdzc(k));
  }
  OCB_X T divv(int i, int j, int k) const {   // cfc
    T s = (Uv(i + 1, j, k) - Uv(i, j, k)) + (Vv(i, j, k) - Vv(i, j - 1, k)) + (Wv(i, j, k + 1) - Wv(i, j, k));
    return s * (X.g.rdx * X.g.rdy * X.rdzc(k));
  }
  OCB_X T divw(int i, int j, int k) const {   // ccf
    T s = (Uw(i + 1, j, k) - Uw(i, j, k)) + (Vw(i, j + 1, k) - Vw(i, j, k)) + (Ww(i, j, k) - Ww(i, j, k - 1));
    return s * (X.g.rdx * X.g.rdy * X.rdzf(k));
  }
};
OCB_EVAL(OCB_DIAG_ADV) {
  AdvFlux<T> A(X, flags);
  T a = X.u(i, j, k) * A.divu(i, j, k) + X.u(i + 1, j, k) * A.divu(i + 1, j, k);
  T b = X.v(i, j, k) * A.divv(i, j, k) + X.v(i, j + 1, k) * A.divv(i, j + 1, k);
  T c = X.w(i, j, k) * A.divw(i, j, k) + X.w(i, j, k + 1) * A.divw(i, j, k + 1);
  return T(0.5) * (a + b + c);
} };

// uᵢ∂ᵢpᶜᶜᶜ  (src/KineticEnergyEquation.jl:304-309)
OCB_EVAL(OCB_DIAG_PR) {
  const HFld<T>& p = X.in.p;
  T pc = p(i, j, k);
  T a = (X.u(i, j, k) * (pc - p(i - 1, j, k)) + X.u(i + 1, j, k) * (p(i + 1, j, k) - pc)) * X.g.rdx;
  T b = (X.v(i, j, k) * (pc - p(i, j - 1, k)) + X.v(i, j + 1, k) * (p(i, j + 1, k) - pc)) * X.g.rdy;
  T c = X.w(i, j, k) * ((pc - p(i, j, k - 1)) * X.rdzf(k)) + X.w(i, j, k + 1) * ((p(i, j, k + 1) - pc) * X.rdzf(k + 1));
  return T(0.5) * (a + b + c);
} };

// uᵢbᵢᶜᶜᶜ  (src/KineticEnergyEquation.jl:362-367): velocity times the buoyancy projected on each axis
OCB_EVAL(OCB_DIAG_WB) {
  const HFld<T>& b = X.in.b;
  const T* gh = X.in.ghat;
  T bc = b(i, j, k);
  T r = T(0);
  if (gh[0] != T(0)) r += gh[0] * (X.u(i, j, k) * (b(i - 1, j, k) + bc) + X.u(i + 1, j, k) * (bc + b(i + 1, j, k)));
  if (gh[1] != T(0)) r += gh[1] * (X.v(i, j, k) * (b(i, j - 1, k) + bc) + X.v(i, j + 1, k) * (bc + b(i, j + 1, k)));
  if (gh[2] != T(0)) r += gh[2] * (X.w(i, j, k) * (b(i, j, k - 1) + bc) + X.w(i, j, k + 1) * (bc + b(i, j, k + 1)));
  return T(0.25) * r;
} };

// viscous_dissipation_rate_ccc  (src/KineticEnergyEquation.jl:427-453), conservative A·δu·τ / V form.
// Per stress location the two symmetric terms are merged:  -(A_a δ_a u_b + A_b δ_b u_a) τ_ab  with
// τ_ab = -2 ν Σ_ab, and  A_a δ_a u_b + A_b δ_b u_a = 2 V_loc Σ_ab,  so each edge contributes 4 ν V_loc Σ_ab².
template <class T> struct EpsEdges {
  const Ctx<T>& X;
  OCB_X EpsEdges(const Ctx<T>& x) : X(x) {}
  OCB_X T e12(int i, int j, int k) const { return X.nu_ffc(i, j, k) * ocb_sq(X.s12(i, j, k)); }              // x (4 V_ffc), V_ffc = dx dy dzc(k)
  OCB_X T e13(int i, int j, int k) const { return X.dzf(k) * (X.nu_fcf(i, j, k) * ocb_sq(X.s13(i, j, k))); }  // x (4 dx dy)
  OCB_X T e23(int i, int j, int k) const { return X.dzf(k) * (X.nu_cff(i, j, k) * ocb_sq(X.s23(i, j, k))); }  // x (4 dx dy)
};
OCB_EVAL(OCB_DIAG_EPS) {
  EpsEdges<T> E(X);
  // diagonal terms: -A δu τ = 2 ν V Σ_aa²  -> after /V: 2 ν Σ_aa²
  T diag = T(2) * X.nu_ccc(i, j, k) * (ocb_sq(X.s11(i, j, k)) + ocb_sq(X.s22(i, j, k)) + ocb_sq(X.s33(i, j, k)));
  T m12 = OCB_MEAN_XY(E.e12, i, j, k);                    // same dzc(k) at the four corners: cancels with 1/V
  T m13 = OCB_MEAN_XZ(E.e13, i, j, k) * X.rdzc(k);
  T m23 = OCB_MEAN_YZ(E.e23, i, j, k) * X.rdzc(k);
  return diag + T(4) * (m12 + m13 + m23);
} };

// strain / rotation sums shared by the moduli (src/FlowDiagnostics.jl:429-441, 575-587)
template <class T> struct SqEdges {
  const Ctx<T>& X;
  OCB_X SqEdges(const Ctx<T>& x) : X(x) {}
  OCB_X T p12(int i, int j, int k) const { return ocb_sq(X.dyu_ffc(i, j, k) + X.dxv_ffc(i, j, k)); }
  OCB_X T p13(int i, int j, int k) const { return ocb_sq(X.dzu_fcf(i, j, k) + X.dxw_fcf(i, j, k)); }
  OCB_X T p23(int i, int j, int k) const { return ocb_sq(X.dzv_cff(i, j, k) + X.dyw_cff(i, j, k)); }
  OCB_X T m12(int i, int j, int k) const { return ocb_sq(X.dyu_ffc(i, j, k) - X.dxv_ffc(i, j, k)); }
  OCB_X T m13(int i, int j, int k) const { return ocb_sq(X.dzu_fcf(i, j, k) - X.dxw_fcf(i, j, k)); }
  OCB_X T m23(int i, int j, int k) const { return ocb_sq(X.dzv_cff(i, j, k) - X.dyw_cff(i, j, k)); }
};
template <class T> OCB_X T ocb_strain_sq(const Ctx<T>& X, int i, int j, int k) {   // S_ij S_ij
  SqEdges<T> E(X);
  T d = ocb_sq(X.s11(i, j, k)) + ocb_sq(X.s22(i, j, k)) + ocb_sq(X.s33(i, j, k));
  T o = OCB_MEAN_XY(E.p12, i, j, k) + OCB_MEAN_XZ(E.p13, i, j, k) + OCB_MEAN_YZ(E.p23, i, j, k);
  return d + T(0.5) * o;   // 2 * (mean of (a+b)^2 / 4)
}
template <class T> OCB_X T ocb_rotation_sq(const Ctx<T>& X, int i, int j, int k) {   // Ω_ij Ω_ij (each pair twice)
  SqEdges<T> E(X);
  T o = OCB_MEAN_XY(E.m12, i, j, k) + OCB_MEAN_XZ(E.m13, i, j, k) + OCB_MEAN_YZ(E.m23, i, j, k);
  return T(0.5) * o;
}
template <class T> OCB_X T ocb_sqrt(T x);
template <> OCB_X double ocb_sqrt<double>(double x) { return sqrt(x); }
template <> OCB_X float ocb_sqrt<float>(float x) { return sqrtf(x); }

// isotropic_viscous_dissipation_rate_ccc  (src/KineticEnergyEquation.jl:497-510), ν at ccc summed over the tuple
OCB_EVAL(OCB_DIAG_EPS_ISO) { return T(2) * X.nu_ccc(i, j, k) * ocb_strain_sq(X, i, j, k); } };
// strain_rate_tensor_modulus_ccc  (src/FlowDiagnostics.jl:431-441)
OCB_EVAL(OCB_DIAG_SMOD) { return ocb_sqrt(ocb_strain_sq(X, i, j, k)); } };
// vorticity_tensor_modulus_ccc  (src/FlowDiagnostics.jl:577-587)
OCB_EVAL(OCB_DIAG_OMOD) { return ocb_sqrt(ocb_rotation_sq(X, i, j, k)); } };
// Q_velocity_gradient_tensor_invariant_ccc  (src/FlowDiagnostics.jl:708-712): (|Ω|² - |S|²)/2 through the roots
OCB_EVAL(OCB_DIAG_Q) {
  T s = ocb_sqrt(ocb_strain_sq(X, i, j, k)), o = ocb_sqrt(ocb_rotation_sq(X, i, j, k));
  return T(0.5) * (o * o - s * s);
} };

// StrainRateTensorKernel{I,J} / VorticityTensorKernel{I,J}  (src/FlowDiagnostics.jl:484-491, 630-634)
OCB_EVAL(OCB_DIAG_S11) { return X.s11(i, j, k); } };
OCB_EVAL(OCB_DIAG_S22) { return X.s22(i, j, k); } };
OCB_EVAL(OCB_DIAG_S33) { return X.s33(i, j, k); } };
OCB_EVAL(OCB_DIAG_S12) { return X.s12(i, j, k); } };
OCB_EVAL(OCB_DIAG_S13) { return X.s13(i, j, k); } };
OCB_EVAL(OCB_DIAG_S23) { return X.s23(i, j, k); } };
OCB_EVAL(OCB_DIAG_O12) { return T(0.5) * (X.dyu_ffc(i, j, k) - X.dxv_ffc(i, j, k)); } };
OCB_EVAL(OCB_DIAG_O13) { return T(0.5) * (X.dzu_fcf(i, j, k) - X.dxw_fcf(i, j, k)); } };
OCB_EVAL(OCB_DIAG_O23) { return T(0.5) * (X.dzv_cff(i, j, k) - X.dyw_cff(i, j, k)); } };

// StressTensorKernel{I,J,Collocated}  (src/FlowDiagnostics.jl:723-733)
OCB_EVAL(OCB_DIAG_T11) { return ocb_sq(X.u(i, j, k)); } };
OCB_EVAL(OCB_DIAG_T22) { return ocb_sq(X.v(i, j, k)); } };
OCB_EVAL(OCB_DIAG_T33) { return ocb_sq(X.w(i, j, k)); } };
OCB_EVAL(OCB_DIAG_T11C) { return ocb_sq(T(0.5) * (X.u(i, j, k) + X.u(i + 1, j, k))); } };
OCB_EVAL(OCB_DIAG_T22C) { return ocb_sq(T(0.5) * (X.v(i, j, k) + X.v(i, j + 1, k))); } };
OCB_EVAL(OCB_DIAG_T33C) { return ocb_sq(T(0.5) * (X.w(i, j, k) + X.w(i, j, k + 1))); } };
OCB_EVAL(OCB_DIAG_T12) { return (T(0.5) * (X.u(i, j - 1, k) + X.u(i, j, k))) * (T(0.5) * (X.v(i - 1, j, k) + X.v(i, j, k))); } };
OCB_EVAL(OCB_DIAG_T13) { return (T(0.5) * (X.u(i, j, k - 1) + X.u(i, j, k))) * (T(0.5) * (X.w(i - 1, j, k) + X.w(i, j, k))); } };
OCB_EVAL(OCB_DIAG_T23) { return (T(0.5) * (X.v(i, j, k - 1) + X.v(i, j, k))) * (T(0.5) * (X.w(i, j - 1, k) + X.w(i, j, k))); } };

// shear_production_rate_{x,y,z}_ccc  (src/TurbulentKineticEnergyEquation.jl:101-117, 163-179, 224-240)
template <class T> struct MeanGrad {   // gradients of the mean flow at their natural points
  const Ctx<T>& X;
  OCB_X MeanGrad(const Ctx<T>& x) : X(x) {}
  OCB_X T dxV(int i, int j, int k) const { return (X.in.V(i, j, k) - X.in.V(i - 1, j, k)) * X.g.rdx; }   // ffc
  OCB_X T dxW(int i, int j, int k) const { return (X.in.W(i, j, k) - X.in.W(i - 1, j, k)) * X.g.rdx; }   // fcf
  OCB_X T dyU(int i, int j, int k) const { return (X.in.U(i, j, k) - X.in.U(i, j - 1, k)) * X.g.rdy; }   // ffc
  OCB_X T dyW(int i, int j, int k) const { return (X.in.W(i, j, k) - X.in.W(i, j - 1, k)) * X.g.rdy; }   // cff
  OCB_X T dzU(int i, int j, int k) const { return (X.in.U(i, j, k) - X.in.U(i, j, k - 1)) * X.rdzf(k); } // fcf
  OCB_X T dzV(int i, int j, int k) const { return (X.in.V(i, j, k) - X.in.V(i, j, k - 1)) * X.rdzf(k); } // cff
};
template <class T> OCB_X T ocb_spx(const Ctx<T>& X, int i, int j, int k) {
  MeanGrad<T> M(X);
  T uc = T(0.5) * (X.u(i, j, k) + X.u(i + 1, j, k));
  T vc = T(0.5) * (X.v(i, j, k) + X.v(i, j + 1, k));
  T wc = T(0.5) * (X.w(i, j, k) + X.w(i, j, k + 1));
  T uu = T(0.5) * (ocb_sq(X.u(i, j, k)) + ocb_sq(X.u(i + 1, j, k)));
  T dxU = (X.in.U(i + 1, j, k) - X.in.U(i, j, k)) * X.g.rdx;
  return -(uu * dxU + (vc * uc) * OCB_MEAN_XY(M.dxV, i, j, k) + (wc * uc) * OCB_MEAN_XZ(M.dxW, i, j, k));
}
template <class T> OCB_X T ocb_spy(const Ctx<T>& X, int i, int j, int k) {
  MeanGrad<T> M(X);
  T uc = T(0.5) * (X.u(i, j, k) + X.u(i + 1, j, k));
  T vc = T(0.5) * (X.v(i, j, k) + X.v(i, j + 1, k));
  T wc = T(0.5) * (X.w(i, j, k) + X.w(i, j, k + 1));
  T vv = T(0.5) * (ocb_sq(X.v(i, j, k)) + ocb_sq(X.v(i, j + 1, k)));
  T dyV = (X.in.V(i, j + 1, k) - X.in.V(i, j, k)) * X.g.rdy;
  return -((uc * vc) * OCB_MEAN_XY(M.dyU, i, j, k) + vv * dyV + (wc * vc) * OCB_MEAN_YZ(M.dyW, i, j, k));
}
template <class T> OCB_X T ocb_spz(const Ctx<T>& X, int i, int j, int k) {
  MeanGrad<T> M(X);
  T uc = T(0.5) * (X.u(i, j, k) + X.u(i + 1, j, k));
  T vc = T(0.5) * (X.v(i, j, k) + X.v(i, j + 1, k));
  T wc = T(0.5) * (X.w(i, j, k) + X.w(i, j, k + 1));
  T ww = T(0.5) * (ocb_sq(X.w(i, j, k)) + ocb_sq(X.w(i, j, k + 1)));
  T dzW = (X.in.W(i, j, k + 1) - X.in.W(i, j, k)) * X.rdzc(k);
  return -((uc * wc) * OCB_MEAN_XZ(M.dzU, i, j, k) + (vc * wc) * OCB_MEAN_YZ(M.dzV, i, j, k) + ww * dzW);
}
OCB_EVAL(OCB_DIAG_SPX) { return ocb_spx(X, i, j, k); } };
OCB_EVAL(OCB_DIAG_SPY) { return ocb_spy(X, i, j, k); } };
OCB_EVAL(OCB_DIAG_SPZ) { return ocb_spz(X, i, j, k); } };
OCB_EVAL(OCB_DIAG_SP) { return ocb_spx(X, i, j, k) + ocb_spy(X, i, j, k) + ocb_spz(X, i, j, k); } };   // :285-289

// tracer_variance_dissipation_rate_ccc  (src/TracerVarianceEquation.jl:146-161): 2/V Σ ℑ(A κ δc²/Δ)
OCB_EVAL(OCB_DIAG_CHI) {
  T cc = X.c(i, j, k);
  T fx0 = X.ka_fcc(i, j, k) * ocb_sq(cc - X.c(i - 1, j, k)), fx1 = X.ka_fcc(i + 1, j, k) * ocb_sq(X.c(i + 1, j, k) - cc);
  T fy0 = X.ka_cfc(i, j, k) * ocb_sq(cc - X.c(i, j - 1, k)), fy1 = X.ka_cfc(i, j + 1, k) * ocb_sq(X.c(i, j + 1, k) - cc);
  T fz0 = X.ka_ccf(i, j, k) * ocb_sq(cc - X.c(i, j, k - 1)) * X.rdzf(k);
  T fz1 = X.ka_ccf(i, j, k + 1) * ocb_sq(X.c(i, j, k + 1) - cc) * X.rdzf(k + 1);
  return (fx0 + fx1) * (X.g.rdx * X.g.rdx) + (fy0 + fy1) * (X.g.rdy * X.g.rdy) + (fz0 + fz1) * X.rdzc(k);
} };

// c∇_dot_qᶜ  (src/TracerVarianceEquation.jl:93-98): 2 c ∇·q with q = -κ ∇c
OCB_EVAL(OCB_DIAG_CDIFF) {
  T cc = X.c(i, j, k);
  T qx0 = X.ka_fcc(i, j, k) * (cc - X.c(i - 1, j, k)), qx1 = X.ka_fcc(i + 1, j, k) * (X.c(i + 1, j, k) - cc);
  T qy0 = X.ka_cfc(i, j, k) * (cc - X.c(i, j - 1, k)), qy1 = X.ka_cfc(i, j + 1, k) * (X.c(i, j + 1, k) - cc);
  T qz0 = X.ka_ccf(i, j, k) * (cc - X.c(i, j, k - 1)) * X.rdzf(k);
  T qz1 = X.ka_ccf(i, j, k + 1) * (X.c(i, j, k + 1) - cc) * X.rdzf(k + 1);
  T div = (qx0 - qx1) * (X.g.rdx * X.g.rdx) + (qy0 - qy1) * (X.g.rdy * X.g.rdy) + (qz0 - qz1) * X.rdzc(k);
  return T(2) * cc * div;
} };

// ---- richardson_number_ccf  (src/FlowDiagnostics.jl:36-68) -----------------------------------------------
template <class T> OCB_X T ocb_uh_norm(const Ctx<T>& X, int i, int j, int k) {   // uₕ_norm_ccc, :48-53
  const T* d = X.in.dir;
  T u0 = X.u(i, j, k), u1 = X.u(i + 1, j, k), v0 = X.v(i, j, k), v1 = X.v(i, j + 1, k), w0 = X.w(i, j, k), w1 = X.w(i, j, k + 1);
  T sq = T(0.5) * ((u0 * u0 + u1 * u1) + (v0 * v0 + v1 * v1) + (w0 * w0 + w1 * w1));
  T along = T(0.5) * ((u0 + u1) * d[0] + (v0 + v1) * d[1] + (w0 + w1) * d[2]);
  T r = sq - along * along;
  return ocb_sqrt(r > T(0) ? r : T(0));
}
OCB_EVAL(OCB_DIAG_RI) {
  const T* d = X.in.dir;
  const HFld<T>& b = X.in.b;
  T rzf = X.rdzf(k);
  T db = (b(i, j, k) - b(i, j, k - 1)) * rzf * d[2];
  T uh0 = ocb_uh_norm(X, i, j, k), uhm = ocb_uh_norm(X, i, j, k - 1);
  T du = (uh0 - uhm) * rzf * d[2];
  if (d[0] != T(0)) {   // tilted gravity: x terms (b gradient averaged to ccf, speed gradient left at level k)
    T bx = T(0.25) * ((b(i + 1, j, k) - b(i - 1, j, k)) + (b(i + 1, j, k - 1) - b(i - 1, j, k - 1))) * X.g.rdx;
    db += bx * d[0];
    du += T(0.5) * (ocb_uh_norm(X, i + 1, j, k) - ocb_uh_norm(X, i - 1, j, k)) * X.g.rdx * d[0];
  }
  if (d[1] != T(0)) {
    T by = T(0.25) * ((b(i, j + 1, k) - b(i, j - 1, k)) + (b(i, j + 1, k - 1) - b(i, j - 1, k - 1))) * X.g.rdy;
    db += by * d[1];
    du += T(0.5) * (ocb_uh_norm(X, i, j + 1, k) - ocb_uh_norm(X, i, j - 1, k)) * X.g.rdy * d[1];
  }
  return db / (du * du);
} };

// ---- vorticity / potential vorticity at fff ----------------------------------------------------------------
template <class T> struct Vort {   // the three relative-vorticity components and ∇b, all at fff
  const Ctx<T>& X;
  OCB_X Vort(const Ctx<T>& x) : X(x) {}
  OCB_X T dwdy(int i, int j, int k) const { return T(0.5) * (X.dyw_cff(i - 1, j, k) + X.dyw_cff(i, j, k)); }
  OCB_X T dvdz(int i, int j, int k) const { return T(0.5) * (X.dzv_cff(i - 1, j, k) + X.dzv_cff(i, j, k)); }
  OCB_X T dudz(int i, int j, int k) const { return T(0.5) * (X.dzu_fcf(i, j - 1, k) + X.dzu_fcf(i, j, k)); }
  OCB_X T dwdx(int i, int j, int k) const { return T(0.5) * (X.dxw_fcf(i, j - 1, k) + X.dxw_fcf(i, j, k)); }
  OCB_X T dvdx(int i, int j, int k) const { return T(0.5) * (X.dxv_ffc(i, j, k - 1) + X.dxv_ffc(i, j, k)); }
  OCB_X T dudy(int i, int j, int k) const { return T(0.5) * (X.dyu_ffc(i, j, k - 1) + X.dyu_ffc(i, j, k)); }
  OCB_X T dbdx(int i, int j, int k) const {
    const HFld<T>& b = X.in.b;
    return T(0.25) * (((b(i, j - 1, k - 1) - b(i - 1, j - 1, k - 1)) + (b(i, j, k - 1) - b(i - 1, j, k - 1))) +
                      ((b(i, j - 1, k) - b(i - 1, j - 1, k)) + (b(i, j, k) - b(i - 1, j, k)))) * X.g.rdx;
  }
  OCB_X T dbdy(int i, int j, int k) const {
    const HFld<T>& b = X.in.b;
    return T(0.25) * (((b(i - 1, j, k - 1) - b(i - 1, j - 1, k - 1)) + (b(i, j, k - 1) - b(i, j - 1, k - 1))) +
                      ((b(i - 1, j, k) - b(i - 1, j - 1, k)) + (b(i, j, k) - b(i, j - 1, k)))) * X.g.rdy;
  }
  OCB_X T dbdz(int i, int j, int k) const {
    const HFld<T>& b = X.in.b;
    return T(0.25) * (((b(i - 1, j - 1, k) - b(i - 1, j - 1, k - 1)) + (b(i, j - 1, k) - b(i, j - 1, k - 1))) +
                      ((b(i - 1, j, k) - b(i - 1, j, k - 1)) + (b(i, j, k) - b(i, j, k - 1)))) * X.rdzf(k);
  }
};
// rossby_number_fff  (src/FlowDiagnostics.jl:124-138); ro_bg = dWdy, dVdz, dUdz, dWdx, dUdy, dVdx
OCB_EVAL(OCB_DIAG_RO) {
  Vort<T> Vt(X);
  const T* bg = X.in.ro_bg; const T* f = X.in.f;
  T wx = (Vt.dwdy(i, j, k) + bg[0]) - (Vt.dvdz(i, j, k) + bg[1]);
  T wy = (Vt.dudz(i, j, k) + bg[2]) - (Vt.dwdx(i, j, k) + bg[3]);
  T wz = (Vt.dvdx(i, j, k) + bg[5]) - (Vt.dudy(i, j, k) + bg[4]);
  return (wx * f[0] + wy * f[1] + wz * f[2]) / (f[0] * f[0] + f[1] * f[1] + f[2] * f[2]);
} };
// ertel_potential_vorticity_fff  (src/FlowDiagnostics.jl:216-233)
OCB_EVAL(OCB_DIAG_EPV) {
  Vort<T> Vt(X);
  const T* f = X.in.f;
  T px = (f[0] + Vt.dwdy(i, j, k) - Vt.dvdz(i, j, k)) * Vt.dbdx(i, j, k);
  T py = (f[1] + Vt.dudz(i, j, k) - Vt.dwdx(i, j, k)) * Vt.dbdy(i, j, k);
  T pz = (f[2] + Vt.dvdx(i, j, k) - Vt.dudy(i, j, k)) * Vt.dbdz(i, j, k);
  return px + py + pz;
} };
// potential_vorticity_in_thermal_wind_fff  (src/FlowDiagnostics.jl:200-214); f = f[2]
OCB_EVAL(OCB_DIAG_TWPV) {
  Vort<T> Vt(X);
  T f = X.in.f[2];
  T barot = (f + Vt.dvdx(i, j, k) - Vt.dudy(i, j, k)) * Vt.dbdz(i, j, k);
  T baroc = -f * (ocb_sq(Vt.dudz(i, j, k)) + ocb_sq(Vt.dvdz(i, j, k)));
  return barot + baroc;
} };
// directional_ertel_potential_vorticity_fff  (src/FlowDiagnostics.jl:358-380)
OCB_EVAL(OCB_DIAG_DEPV) {
  Vort<T> Vt(X);
  const T* d = X.in.dir;
  T ox = Vt.dwdy(i, j, k) - Vt.dvdz(i, j, k), oy = Vt.dudz(i, j, k) - Vt.dwdx(i, j, k), oz = Vt.dvdx(i, j, k) - Vt.dudy(i, j, k);
  T od = ox * d[0] + oy * d[1] + oz * d[2];
  T bd = Vt.dbdx(i, j, k) * d[0] + Vt.dbdy(i, j, k) * d[1] + Vt.dbdz(i, j, k) * d[2];
  return (X.in.f_dir + od) * bd;
} };

// Oceananigans viscous_flux_uᵢxⱼ = -2 ν Σᵢⱼ at ccc / ffc / fcf / cff (filtered by
// src/FilteredKineticEnergyEquation.jl:377-387)
OCB_EVAL(OCB_DIAG_VF11) { return T(-2) * X.nu_ccc(i, j, k) * X.s11(i, j, k); } };
OCB_EVAL(OCB_DIAG_VF22) { return T(-2) * X.nu_ccc(i, j, k) * X.s22(i, j, k); } };
OCB_EVAL(OCB_DIAG_VF33) { return T(-2) * X.nu_ccc(i, j, k) * X.s33(i, j, k); } };
OCB_EVAL(OCB_DIAG_VF12) { return T(-2) * X.nu_ffc(i, j, k) * X.s12(i, j, k); } };
OCB_EVAL(OCB_DIAG_VF13) { return T(-2) * X.nu_fcf(i, j, k) * X.s13(i, j, k); } };
OCB_EVAL(OCB_DIAG_VF23) { return T(-2) * X.nu_cff(i, j, k) * X.s23(i, j, k); } };

// filtered_dissipation_rate_ccc  (src/FilteredKineticEnergyEquation.jl:275-299): u, v, w are the
// filtered velocities, aux[0..5] the filtered fluxes τ̄11, τ̄22, τ̄33, τ̄12(=τ̄21), τ̄13(=τ̄31), τ̄23(=τ̄32).
template <class T> struct FepsEdges {
  const Ctx<T>& X;
  OCB_X FepsEdges(const Ctx<T>& x) : X(x) {}
  // -(Ay δy ū + Ax δx v̄) τ̄12 = -2 V_ffc S̄12 τ̄12 ; the common dx dy factor is applied by the caller
  OCB_X T e12(int i, int j, int k) const { return X.s12(i, j, k) * X.in.aux[3](i, j, k); }
  OCB_X T e13(int i, int j, int k) const { return X.dzf(k) * (X.s13(i, j, k) * X.in.aux[4](i, j, k)); }
  OCB_X T e23(int i, int j, int k) const { return X.dzf(k) * (X.s23(i, j, k) * X.in.aux[5](i, j, k)); }
};
template <class T> OCB_X T ocb_feps(const Ctx<T>& X, int i, int j, int k) {
  FepsEdges<T> E(X);
  T diag = X.s11(i, j, k) * X.in.aux[0](i, j, k) + X.s22(i, j, k) * X.in.aux[1](i, j, k) + X.s33(i, j, k) * X.in.aux[2](i, j, k);
  T m12 = OCB_MEAN_XY(E.e12, i, j, k);
  T m13 = OCB_MEAN_XZ(E.e13, i, j, k) * X.rdzc(k);
  T m23 = OCB_MEAN_YZ(E.e23, i, j, k) * X.rdzc(k);
  return -(diag + T(2) * (m12 + m13 + m23));
}
OCB_EVAL(OCB_DIAG_FEPS) { return ocb_feps(X, i, j, k); } };
// subfilter_ke_dissipation_rate_ccc  (src/SubFilterKineticEnergyEquation.jl:92): filter(ε) - εˡ
OCB_EVAL(OCB_DIAG_SFEPS) { return X.in.aux[6](i, j, k) - ocb_feps(X, i, j, k); } };
// subfilter_kinetic_energy_ccc  (src/SubFilterKineticEnergyEquation.jl:29): filter(K) - K(ū)
OCB_EVAL(OCB_DIAG_SFKE) { return X.in.aux[6](i, j, k) - ocb_ke(X, i, j, k); } };

// cross_scale_ke_flux_ccc  (src/FilteredKineticEnergyEquation.jl:176-188): Π = -Σ wᵢⱼ ⟨τᵢⱼ⟩ᶜ ⟨S̄ᵢⱼ⟩ᶜ with
// τᵢⱼ = aux_ij - ūᵢūⱼ at its staggered location, every factor brought to ccc by two-point means.
template <class T> struct PiEdges {
  const Ctx<T>& X;
  OCB_X PiEdges(const Ctx<T>& x) : X(x) {}
  OCB_X T t11(int i, int j, int k) const { return X.in.aux[0](i, j, k) - ocb_sq(X.u(i, j, k)); }   // fcc
  OCB_X T t22(int i, int j, int k) const { return X.in.aux[1](i, j, k) - ocb_sq(X.v(i, j, k)); }   // cfc
  OCB_X T t33(int i, int j, int k) const { return X.in.aux[2](i, j, k) - ocb_sq(X.w(i, j, k)); }   // ccf
  OCB_X T t12(int i, int j, int k) const { return X.in.aux[3](i, j, k) - Eval<T, OCB_DIAG_T12>::at(X, 0, i, j, k); }
  OCB_X T t13(int i, int j, int k) const { return X.in.aux[4](i, j, k) - Eval<T, OCB_DIAG_T13>::at(X, 0, i, j, k); }
  OCB_X T t23(int i, int j, int k) const { return X.in.aux[5](i, j, k) - Eval<T, OCB_DIAG_T23>::at(X, 0, i, j, k); }
  OCB_X T s12(int i, int j, int k) const { return X.s12(i, j, k); }
  OCB_X T s13(int i, int j, int k) const { return X.s13(i, j, k); }
  OCB_X T s23(int i, int j, int k) const { return X.s23(i, j, k); }
};
OCB_EVAL(OCB_DIAG_PI) {
  PiEdges<T> E(X);
  const bool d1 = flags & OCB_REQ_TENSOR_DIM1, d2 = flags & OCB_REQ_TENSOR_DIM2, d3 = flags & OCB_REQ_TENSOR_DIM3;
  const bool col = flags & OCB_REQ_COLLOCATED;
  T s = T(0);
  if (d1) s += (col ? X.in.aux[0](i, j, k) - Eval<T, OCB_DIAG_T11C>::at(X, 0, i, j, k) : T(0.5) * (E.t11(i, j, k) + E.t11(i + 1, j, k))) * X.s11(i, j, k);
  if (d2) s += (col ? X.in.aux[1](i, j, k) - Eval<T, OCB_DIAG_T22C>::at(X, 0, i, j, k) : T(0.5) * (E.t22(i, j, k) + E.t22(i, j + 1, k))) * X.s22(i, j, k);
  if (d3) s += (col ? X.in.aux[2](i, j, k) - Eval<T, OCB_DIAG_T33C>::at(X, 0, i, j, k) : T(0.5) * (E.t33(i, j, k) + E.t33(i, j, k + 1))) * X.s33(i, j, k);
  if (d1 && d2) s += T(2) * OCB_MEAN_XY(E.t12, i, j, k) * OCB_MEAN_XY(E.s12, i, j, k);
  if (d1 && d3) s += T(2) * OCB_MEAN_XZ(E.t13, i, j, k) * OCB_MEAN_XZ(E.s13, i, j, k);
  if (d2 && d3) s += T(2) * OCB_MEAN_YZ(E.t23, i, j, k) * OCB_MEAN_YZ(E.s23, i, j, k);
  return -s;
} };

// minus_bz✶_ccc  (src/BackgroundPotentialEnergyEquation.jl:852)
OCB_EVAL(OCB_DIAG_BPE) { return -X.in.b(i, j, k) * X.in.aux[6](i, j, k); } };
// local_ape_ccc  (src/AvailablePotentialEnergyEquation.jl:42-45): Ψδ - b (z - z✶)
// (the parcel's own height is the grid's Zᶜᶜᶜ on the model grid, a carried field -- aux[3], flagged -- on a VerticalSort column, :43)
OCB_EVAL(OCB_DIAG_APE) {
  const T z = (flags & OCB_REQ_CARRIED_HEIGHT) ? X.in.aux[3](i, j, k) : OCB_LD(X.g.zc + k);
  return X.in.aux[5](i, j, k) - X.in.b(i, j, k) * (z - X.in.aux[6](i, j, k));
} };

// upsilon_ccc  (src/AvailablePotentialEnergyEquation.jl:126): Υ = z✶ - z, with aux[6] = z✶
OCB_EVAL(OCB_DIAG_UPSILON) { return X.in.aux[6](i, j, k) - OCB_LD(X.g.zc + k); } };
// ape_dissipation_rate_ccc  (src/AvailablePotentialEnergyEquation.jl:202-215): (1/V) Σ ℑ(-A δΥ q) with q = -κ ∂c, Υ = aux[4]
// -- the discretisation of tracer_variance_dissipation_rate_ccc with one δc replaced by δΥ, and half of it
OCB_EVAL(OCB_DIAG_APE_EPS) {
  const HFld<T>& Y = X.in.aux[4];       // (aux[5], aux[6] stay free for Ψδ and z✶: the whole APE family shares a sweep)
  T cc = X.c(i, j, k), yc = Y(i, j, k);
  T fx0 = X.ka_fcc(i, j, k) * ((cc - X.c(i - 1, j, k)) * (yc - Y(i - 1, j, k)));
  T fx1 = X.ka_fcc(i + 1, j, k) * ((X.c(i + 1, j, k) - cc) * (Y(i + 1, j, k) - yc));
  T fy0 = X.ka_cfc(i, j, k) * ((cc - X.c(i, j - 1, k)) * (yc - Y(i, j - 1, k)));
  T fy1 = X.ka_cfc(i, j + 1, k) * ((X.c(i, j + 1, k) - cc) * (Y(i, j + 1, k) - yc));
  T fz0 = X.ka_ccf(i, j, k) * ((cc - X.c(i, j, k - 1)) * (yc - Y(i, j, k - 1))) * X.rdzf(k);
  T fz1 = X.ka_ccf(i, j, k + 1) * ((X.c(i, j, k + 1) - cc) * (Y(i, j, k + 1) - yc)) * X.rdzf(k + 1);
  return T(0.5) * ((fx0 + fx1) * (X.g.rdx * X.g.rdx) + (fy0 + fy1) * (X.g.rdy * X.g.rdy) + (fz0 + fz1) * X.rdzc(k));
} };

// uᵢ∂ⱼ_τᵢⱼᶜᶜᶜ  (src/KineticEnergyEquation.jl:191-202): each velocity times the divergence of its viscous fluxes at the
// velocity point (Oceananigans ∂ⱼ_τ₁ⱼ = 1/Vᶠᶜᶜ [δxᶠ(Ax τ₁₁) + δyᶜ(Ay τ₁₂) + δzᶜ(Az τ₁₃)] and likewise for v, w; on a grid
// with regular x and y the areas and the volume reduce to the spacings below), interpolated to the cell centre.
template <class T> struct StressDiv {
  const Ctx<T>& X;
  OCB_X StressDiv(const Ctx<T>& x) : X(x) {}
  OCB_X T t11(int i, int j, int k) const { return T(-2) * X.nu_ccc(i, j, k) * X.s11(i, j, k); }
  OCB_X T t22(int i, int j, int k) const { return T(-2) * X.nu_ccc(i, j, k) * X.s22(i, j, k); }
  OCB_X T t33(int i, int j, int k) const { return T(-2) * X.nu_ccc(i, j, k) * X.s33(i, j, k); }
  OCB_X T t12(int i, int j, int k) const { return T(-2) * X.nu_ffc(i, j, k) * X.s12(i, j, k); }
  OCB_X T t13(int i, int j, int k) const { return T(-2) * X.nu_fcf(i, j, k) * X.s13(i, j, k); }
  OCB_X T t23(int i, int j, int k) const { return T(-2) * X.nu_cff(i, j, k) * X.s23(i, j, k); }
  OCB_X T d1(int i, int j, int k) const {   // fcc
    return (t11(i, j, k) - t11(i - 1, j, k)) * X.g.rdx + (t12(i, j + 1, k) - t12(i, j, k)) * X.g.rdy + (t13(i, j, k + 1) - t13(i, j, k)) * X.rdzc(k);
  }
  OCB_X T d2(int i, int j, int k) const {   // cfc
    return (t12(i + 1, j, k) - t12(i, j, k)) * X.g.rdx + (t22(i, j, k) - t22(i, j - 1, k)) * X.g.rdy + (t23(i, j, k + 1) - t23(i, j, k)) * X.rdzc(k);
  }
  OCB_X T d3(int i, int j, int k) const {   // ccf
    return (t13(i + 1, j, k) - t13(i, j, k)) * X.g.rdx + (t23(i, j + 1, k) - t23(i, j, k)) * X.g.rdy + (t33(i, j, k) - t33(i, j, k - 1)) * X.rdzf(k);
  }
};
OCB_EVAL(OCB_DIAG_KE_STRESS) {
  StressDiv<T> D(X);
  T a = X.u(i, j, k) * D.d1(i, j, k) + X.u(i + 1, j, k) * D.d1(i + 1, j, k);
  T b = X.v(i, j, k) * D.d2(i, j, k) + X.v(i, j + 1, k) * D.d2(i, j + 1, k);
  T c = X.w(i, j, k) * D.d3(i, j, k) + X.w(i, j, k + 1) * D.d3(i, j, k + 1);
  return T(0.5) * (a + b + c);
} };

// uᵢFᵤᵢᶜᶜᶜ  (src/KineticEnergyEquation.jl:251-260): each velocity times its forcing at the velocity point, interpolated to
// the cell centre.  The forcing functors are the host layer's business (Oceananigans evaluates them); here aux[0], aux[1],
// aux[2] hold Fᵘ (fcc), Fᵛ (cfc), Fʷ (ccf) as arrays, numbers or ZeroField.
OCB_EVAL(OCB_DIAG_KE_FORCING) {
  const HFld<T>& Fu = X.in.aux[0]; const HFld<T>& Fv = X.in.aux[1]; const HFld<T>& Fw = X.in.aux[2];
  T a = X.u(i, j, k) * Fu(i, j, k) + X.u(i + 1, j, k) * Fu(i + 1, j, k);
  T b = X.v(i, j, k) * Fv(i, j, k) + X.v(i, j + 1, k) * Fv(i, j + 1, k);
  T c = X.w(i, j, k) * Fw(i, j, k) + X.w(i, j, k + 1) * Fw(i, j, k + 1);
  return T(0.5) * (a + b + c);
} };

// TracerEquation.jl:70-327: div_Uc, ∇_dot_qᶜ, the forcing, the whole tracer tendency; and the three diffusive fluxes.
// Advective fluxes Ax u ℑx c, Ay v ℑy c, Az w ℑz c with the Centered(order = 2 | 4) symmetric interpolation of the tracer.
template <class T> struct TracerTerms {
  const Ctx<T>& X;
  AdvFlux<T> A;
  OCB_X TracerTerms(const Ctx<T>& x, int flags = 0) : X(x), A(x, flags) {}
  OCB_X T qx(int i, int j, int k) const { return -X.ka_fcc(i, j, k) * ((X.in.b(i, j, k) - X.in.b(i - 1, j, k)) * X.g.rdx); }
  OCB_X T qy(int i, int j, int k) const { return -X.ka_cfc(i, j, k) * ((X.in.b(i, j, k) - X.in.b(i, j - 1, k)) * X.g.rdy); }
  OCB_X T qz(int i, int j, int k) const { return -X.ka_ccf(i, j, k) * ((X.in.b(i, j, k) - X.in.b(i, j, k - 1)) * X.rdzf(k)); }
  // twice the tracer interpolated to the x / y / z face m (the factor 1/2 of the two-point mean is applied by `adv`)
  OCB_X T cx2(int i, int j, int k) const { const HFld<T>& c = X.in.b; const bool w4 = A.wide(0, i, true);
    return T(2) * A.mix(w4, w4 ? c(i - 2, j, k) : T(0), c(i - 1, j, k), c(i, j, k), w4 ? c(i + 1, j, k) : T(0)); }
  OCB_X T cy2(int i, int j, int k) const { const HFld<T>& c = X.in.b; const bool w4 = A.wide(1, j, true);
    return T(2) * A.mix(w4, w4 ? c(i, j - 2, k) : T(0), c(i, j - 1, k), c(i, j, k), w4 ? c(i, j + 1, k) : T(0)); }
  OCB_X T cz2(int i, int j, int k) const { const HFld<T>& c = X.in.b; const bool w4 = A.wide(2, k, true);
    return T(2) * A.mix(w4, w4 ? c(i, j, k - 2) : T(0), c(i, j, k - 1), c(i, j, k), w4 ? c(i, j, k + 1) : T(0)); }
  OCB_X T adv(int i, int j, int k) const {
    const T ax = (X.u(i + 1, j, k) * cx2(i + 1, j, k) - X.u(i, j, k) * cx2(i, j, k)) * (T(0.5) * X.g.rdx);
    const T ay = (X.v(i, j + 1, k) * cy2(i, j + 1, k) - X.v(i, j, k) * cy2(i, j, k)) * (T(0.5) * X.g.rdy);
    const T az = (X.w(i, j, k + 1) * cz2(i, j, k + 1) - X.w(i, j, k) * cz2(i, j, k)) * (T(0.5) * X.rdzc(k));
    return (ax + ay) + az;
  }
  OCB_X T divq(int i, int j, int k) const {
    return (qx(i + 1, j, k) - qx(i, j, k)) * X.g.rdx + (qy(i, j + 1, k) - qy(i, j, k)) * X.g.rdy + (qz(i, j, k + 1) - qz(i, j, k)) * X.rdzc(k);
  }
};
// c∂ₜcᶜᶜᶜ  (src/TracerVarianceEquation.jl:18-39): 2c x Oceananigans' NonhydrostaticModel tracer_tendency for a model without
// background fields, biogeochemistry or immersed boundaries:  Gc = -div_Uc - ∇·q + F, with the Centered advective fluxes
// Ax u ℑx c, Ay v ℑy c, Az w ℑz c  (div_Uc = 1/V [δx + δy + δz] of them), the closure's diffusive fluxes as in c∇_dot_qᶜ above, and
// the tracer forcing materialised at the cell centres in aux[0].
OCB_EVAL(OCB_DIAG_C_TENDENCY) {
  TracerTerms<T> C(X, flags);
  const HFld<T>& c = X.in.b;
  T cc = c(i, j, k);
  T qx0 = X.ka_fcc(i, j, k) * (cc - c(i - 1, j, k)), qx1 = X.ka_fcc(i + 1, j, k) * (c(i + 1, j, k) - cc);
  T qy0 = X.ka_cfc(i, j, k) * (cc - c(i, j - 1, k)), qy1 = X.ka_cfc(i, j + 1, k) * (c(i, j + 1, k) - cc);
  T qz0 = X.ka_ccf(i, j, k) * (cc - c(i, j, k - 1)) * X.rdzf(k);
  T qz1 = X.ka_ccf(i, j, k + 1) * (c(i, j, k + 1) - cc) * X.rdzf(k + 1);
  T divq = (qx0 - qx1) * (X.g.rdx * X.g.rdx) + (qy0 - qy1) * (X.g.rdy * X.g.rdy) + (qz0 - qz1) * X.rdzc(k);
  return T(2) * cc * (X.in.aux[0](i, j, k) - C.adv(i, j, k) - divq);
} };

// uᵢGᵢᶜᶜᶜ  (src/KineticEnergyEquation.jl:74-95): each velocity times Oceananigans' NonhydrostaticModel velocity tendency at the
// velocity point, interpolated to the cell centre.  For a model without background fields, immersed boundaries or Stokes drift
//   Gu = -div_𝐯u + x_dot_g_b - x_f_cross_U - ∂x pHY′ - ∂ⱼ_τ₁ⱼ + Fu      (v alike)
//   Gw = -div_𝐯w + [no pHY′: z_dot_g_b] - z_f_cross_U - ∂ⱼ_τ₃ⱼ + Fw
// [OCN-mem: Models/NonhydrostaticModels/velocity_and_tracer_tendencies.jl; Coriolis/constant_cartesian_coriolis.jl -- FPlane is
//  f = (0, 0, f)], Centered second-order advection.  The nonhydrostatic pressure gradient is not part of G (:100-103).
template <class T> struct Tendency {
  const Ctx<T>& X;
  AdvFlux<T> A;
  StressDiv<T> D;
  bool w_buoyancy;
  OCB_X Tendency(const Ctx<T>& x, int flags) : X(x), A(x, flags), D(x), w_buoyancy(!(flags & OCB_REQ_HYDROSTATIC_SPLIT)) {}
  // x_dot_g_b, y_dot_g_b, z_dot_g_b  [OCN-mem BuoyancyFormulations/buoyancy_force.jl]
  OCB_X T bx(int i, int j, int k) const { const SweepIn<T>& I = X.in; return I.ghat[0] != T(0) ? I.ghat[0] * (T(0.5) * (I.b(i - 1, j, k) + I.b(i, j, k))) : T(0); }
  OCB_X T by(int i, int j, int k) const { const SweepIn<T>& I = X.in; return I.ghat[1] != T(0) ? I.ghat[1] * (T(0.5) * (I.b(i, j - 1, k) + I.b(i, j, k))) : T(0); }
  OCB_X T bz(int i, int j, int k) const { const SweepIn<T>& I = X.in; return I.ghat[2] != T(0) ? I.ghat[2] * (T(0.5) * (I.b(i, j, k - 1) + I.b(i, j, k))) : T(0); }
  // x_f_cross_U, y_f_cross_U, z_f_cross_U
  OCB_X T cx(int i, int j, int k) const {
    const SweepIn<T>& I = X.in;
    T cor = T(0);
    if (I.f[1] != T(0)) cor += I.f[1] * (T(0.25) * ((X.w(i - 1, j, k) + X.w(i, j, k)) + (X.w(i - 1, j, k + 1) + X.w(i, j, k + 1))));
    if (I.f[2] != T(0)) cor -= I.f[2] * (T(0.25) * ((X.v(i - 1, j, k) + X.v(i, j, k)) + (X.v(i - 1, j + 1, k) + X.v(i, j + 1, k))));
    return cor;
  }
  OCB_X T cy(int i, int j, int k) const {
    const SweepIn<T>& I = X.in;
    T cor = T(0);
    if (I.f[2] != T(0)) cor += I.f[2] * (T(0.25) * ((X.u(i, j - 1, k) + X.u(i + 1, j - 1, k)) + (X.u(i, j, k) + X.u(i + 1, j, k))));
    if (I.f[0] != T(0)) cor -= I.f[0] * (T(0.25) * ((X.w(i, j - 1, k) + X.w(i, j, k)) + (X.w(i, j - 1, k + 1) + X.w(i, j, k + 1))));
    return cor;
  }
  OCB_X T cz(int i, int j, int k) const {
    const SweepIn<T>& I = X.in;
    T cor = T(0);
    if (I.f[0] != T(0)) cor += I.f[0] * (T(0.25) * ((X.v(i, j, k - 1) + X.v(i, j + 1, k - 1)) + (X.v(i, j, k) + X.v(i, j + 1, k))));
    if (I.f[1] != T(0)) cor -= I.f[1] * (T(0.25) * ((X.u(i, j, k - 1) + X.u(i + 1, j, k - 1)) + (X.u(i, j, k) + X.u(i + 1, j, k))));
    return cor;
  }
  // hydrostatic_pressure_gradient_x / _y of the pressure in slot p
  OCB_X T px(int i, int j, int k) const { return (X.in.p(i, j, k) - X.in.p(i - 1, j, k)) * X.g.rdx; }
  OCB_X T py(int i, int j, int k) const { return (X.in.p(i, j, k) - X.in.p(i, j - 1, k)) * X.g.rdy; }
  // the whole tendency reads slot p only when the request says the model carries a hydrostatic pressure anomaly there (siblings
  // of the same sweep may keep the nonhydrostatic pressure in that slot: uᵢ∂ᵢp)
  OCB_X T gu(int i, int j, int k) const { return ((X.in.aux[0](i, j, k) - A.divu(i, j, k) - D.d1(i, j, k) - (w_buoyancy ? T(0) : px(i, j, k))) + bx(i, j, k)) - cx(i, j, k); }   // fcc
  OCB_X T gv(int i, int j, int k) const { return ((X.in.aux[1](i, j, k) - A.divv(i, j, k) - D.d2(i, j, k) - (w_buoyancy ? T(0) : py(i, j, k))) + by(i, j, k)) - cy(i, j, k); }   // cfc
  OCB_X T gw(int i, int j, int k) const {                                                                                                                      // ccf
    T r = X.in.aux[2](i, j, k) - A.divw(i, j, k) - D.d3(i, j, k);
    if (w_buoyancy) r += bz(i, j, k);
    return r - cz(i, j, k);
  }
};
OCB_EVAL(OCB_DIAG_KE_TENDENCY) {
  Tendency<T> G(X, flags);
  T a = X.u(i, j, k) * G.gu(i, j, k) + X.u(i + 1, j, k) * G.gu(i + 1, j, k);
  T b = X.v(i, j, k) * G.gv(i, j, k) + X.v(i, j + 1, k) * G.gv(i, j + 1, k);
  T c = X.w(i, j, k) * G.gw(i, j, k) + X.w(i, j, k + 1) * G.gw(i, j, k + 1);
  return T(0.5) * (a + b + c);
} };

// The tendency terms one by one (UMomentumEquation.jl:99-504 and its v / w twins): the raw Oceananigans kernels, unsigned
OCB_EVAL(OCB_DIAG_U_TERM) {
  Tendency<T> G(X, flags);
  switch ((flags >> OCB_REQ_TERM_SHIFT) & 15) {
    case OCB_TERM_ADVECTION: return G.A.divu(i, j, k);
    case OCB_TERM_BUOYANCY: return G.bx(i, j, k);
    case OCB_TERM_CORIOLIS: return G.cx(i, j, k);
    case OCB_TERM_PRESSURE: return G.px(i, j, k);
    case OCB_TERM_DIFFUSION: return G.D.d1(i, j, k);
    case OCB_TERM_FORCING: return X.in.aux[0](i, j, k);
    default: return G.gu(i, j, k);
  }
} };
OCB_EVAL(OCB_DIAG_V_TERM) {
  Tendency<T> G(X, flags);
  switch ((flags >> OCB_REQ_TERM_SHIFT) & 15) {
    case OCB_TERM_ADVECTION: return G.A.divv(i, j, k);
    case OCB_TERM_BUOYANCY: return G.by(i, j, k);
    case OCB_TERM_CORIOLIS: return G.cy(i, j, k);
    case OCB_TERM_PRESSURE: return G.py(i, j, k);
    case OCB_TERM_DIFFUSION: return G.D.d2(i, j, k);
    case OCB_TERM_FORCING: return X.in.aux[1](i, j, k);
    default: return G.gv(i, j, k);
  }
} };
OCB_EVAL(OCB_DIAG_W_TERM) {
  Tendency<T> G(X, flags);
  switch ((flags >> OCB_REQ_TERM_SHIFT) & 15) {
    case OCB_TERM_ADVECTION: return G.A.divw(i, j, k);
    case OCB_TERM_BUOYANCY: return G.bz(i, j, k);
    case OCB_TERM_CORIOLIS: return G.cz(i, j, k);
    case OCB_TERM_PRESSURE: return T(0);                 // the vertical momentum equation carries no hydrostatic pressure gradient
    case OCB_TERM_DIFFUSION: return G.D.d3(i, j, k);
    case OCB_TERM_FORCING: return X.in.aux[2](i, j, k);
    default: return G.gw(i, j, k);
  }
} };
// (TracerTerms: defined above, next to c∂ₜcᶜᶜᶜ)
OCB_EVAL(OCB_DIAG_C_TERM) {
  TracerTerms<T> C(X, flags);
  switch ((flags >> OCB_REQ_TERM_SHIFT) & 15) {
    case OCB_TERM_ADVECTION: return C.adv(i, j, k);
    case OCB_TERM_DIFFUSION: return C.divq(i, j, k);
    case OCB_TERM_FORCING: return X.in.aux[0](i, j, k);
    default: return (X.in.aux[0](i, j, k) - C.adv(i, j, k)) - C.divq(i, j, k);
  }
} };
OCB_EVAL(OCB_DIAG_QX) { return TracerTerms<T>(X).qx(i, j, k); } };
OCB_EVAL(OCB_DIAG_QY) { return TracerTerms<T>(X).qy(i, j, k); } };
OCB_EVAL(OCB_DIAG_QZ) { return TracerTerms<T>(X).qz(i, j, k); } };

// Smagorinsky-Lilly eddy viscosity at ccc (Oceananigans' compute_diffusivities! for SmagorinskyLilly,
// the step before the path) [OCN-mem: TurbulenceClosures/turbulence_closure_implementations/smagorinsky_lilly.jl]:
//   νₑ = ς (C Δᶠ)² √(2 Σ²),  Σ² = ΣᵢⱼΣᵢⱼᶜᶜᶜ (the square of strain_rate_tensor_modulus_ccc, src/FlowDiagnostics.jl:431-441),
//   Δᶠ = ∛(Δx Δy Δz),  ς = Σ² == 0 ? 0 : √(max(0, 1 - Cb N²⁺/(Pr Σ²))),  N²⁺ = max(0, ℑzᵃᵃᶜ ∂zᶜᶜᶠ b)
template <class T> OCB_X T ocb_cbrt(T x);
template <> OCB_X double ocb_cbrt<double>(double x) { return cbrt(x); }
template <> OCB_X float ocb_cbrt<float>(float x) { return cbrtf(x); }
OCB_EVAL(OCB_DIAG_NU_SMAG) {
  const SweepIn<T>& I = X.in;
  const T S2 = ocb_strain_sq(X, i, j, k);
  const T C = I.les[0], Cb = I.les[1], Pr = I.les[2];
  const T delta = ocb_cbrt(X.g.dx * X.g.dy * X.dzc(k));
  T sig = T(1);
  if (Cb != T(0)) {
    const T bc = I.b(i, j, k);
    T N2 = T(0.5) * ((bc - I.b(i, j, k - 1)) * X.rdzf(k) + (I.b(i, j, k + 1) - bc) * X.rdzf(k + 1));
    N2 = N2 > T(0) ? N2 : T(0);
    const T s2 = T(1) - Cb * N2 / (Pr * S2);
    sig = S2 == T(0) ? T(0) : ocb_sqrt(s2 > T(0) ? s2 : T(0));
  }
  return sig * ocb_sq(C * delta) * ocb_sqrt(T(2) * S2);
} };

// ---- dispatch -------------------------------------------------------------------------------------------------
#define OCB_FOR_EACH_DIAG(M) \
  M(OCB_DIAG_KE) M(OCB_DIAG_ADV) M(OCB_DIAG_PR) M(OCB_DIAG_WB) M(OCB_DIAG_EPS) M(OCB_DIAG_EPS_ISO) M(OCB_DIAG_TKE) \
  M(OCB_DIAG_SPX) M(OCB_DIAG_SPY) M(OCB_DIAG_SPZ) M(OCB_DIAG_SP) M(OCB_DIAG_CHI) M(OCB_DIAG_CDIFF) M(OCB_DIAG_SMOD) \
  M(OCB_DIAG_OMOD) M(OCB_DIAG_Q) M(OCB_DIAG_S11) M(OCB_DIAG_S22) M(OCB_DIAG_S33) M(OCB_DIAG_S12) M(OCB_DIAG_S13) \
  M(OCB_DIAG_S23) M(OCB_DIAG_O12) M(OCB_DIAG_O13) M(OCB_DIAG_O23) M(OCB_DIAG_T11) M(OCB_DIAG_T22) M(OCB_DIAG_T33) \
  M(OCB_DIAG_T11C) M(OCB_DIAG_T22C) M(OCB_DIAG_T33C) M(OCB_DIAG_T12) M(OCB_DIAG_T13) M(OCB_DIAG_T23) M(OCB_DIAG_RI) \
  M(OCB_DIAG_RO) M(OCB_DIAG_EPV) M(OCB_DIAG_TWPV) M(OCB_DIAG_DEPV) M(OCB_DIAG_VF11) M(OCB_DIAG_VF22) M(OCB_DIAG_VF33) \
  M(OCB_DIAG_VF12) M(OCB_DIAG_VF13) M(OCB_DIAG_VF23) M(OCB_DIAG_FEPS) M(OCB_DIAG_SFEPS) M(OCB_DIAG_SFKE) M(OCB_DIAG_PI) \
  M(OCB_DIAG_BPE) M(OCB_DIAG_APE) M(OCB_DIAG_UPSILON) M(OCB_DIAG_APE_EPS) M(OCB_DIAG_KE_STRESS) M(OCB_DIAG_KE_FORCING) M(OCB_DIAG_C_TENDENCY) \
  M(OCB_DIAG_KE_TENDENCY) M(OCB_DIAG_NU_SMAG) M(OCB_DIAG_U_TERM) M(OCB_DIAG_V_TERM) M(OCB_DIAG_W_TERM) M(OCB_DIAG_C_TERM) \
  M(OCB_DIAG_QX) M(OCB_DIAG_QY) M(OCB_DIAG_QZ)

// diagnostics whose operands may be lazy perturbations (u - U, c - C); the flag is rejected for the others
OCB_X constexpr bool ocb_diag_takes_perturbation(int d) {
  return d == OCB_DIAG_KE || d == OCB_DIAG_EPS || d == OCB_DIAG_EPS_ISO || d == OCB_DIAG_SPX || d == OCB_DIAG_SPY || d == OCB_DIAG_SPZ ||
         d == OCB_DIAG_SP || d == OCB_DIAG_CHI || d == OCB_DIAG_CDIFF || d == OCB_DIAG_SMOD || d == OCB_DIAG_OMOD || d == OCB_DIAG_Q;
}

// ---- cell-relative addressing ------------------------------------------------------------------------------------
// The evaluators address their operands as f(i + di, j + dj, k + dk).  Evaluating them at the literal origin
// (0, 0, 0) of a context whose field pointers and z tables have been shifted to the current cell turns every such
// access into  base[const], with the constants (±sx, ±sy, ±sz and their sums) folded and shared across accesses --
// instead of three 64-bit multiplies per access (ncu: 4000 integer instructions per cell before this change).
template <class T>
OCB_X void ocb_shift_field(HFld<T>& f, long long i, long long j, long long k) {
  if (f.p) f.p += i * f.sx + j * f.sy + k * f.sz;
}
template <class T>
OCB_X void ocb_shift_inputs(SweepIn<T>& L, DGrid<T>& G, int i, int j, int k) {
  ocb_shift_field(L.u, i, j, k); ocb_shift_field(L.v, i, j, k); ocb_shift_field(L.w, i, j, k);
  ocb_shift_field(L.b, i, j, k); ocb_shift_field(L.p, i, j, k);
  ocb_shift_field(L.U, i, j, k); ocb_shift_field(L.V, i, j, k); ocb_shift_field(L.W, i, j, k); ocb_shift_field(L.B, i, j, k);
  for (int a = 0; a < OCB_N_AUX; ++a) ocb_shift_field(L.aux[a], i, j, k);
  for (int a = 0; a < OCB_MAX_CLOSURE_FIELDS; ++a) { ocb_shift_field(L.nu_f[a], i, j, k); ocb_shift_field(L.ka_f[a], i, j, k); }
  G.dzc += k; G.dzf += k; G.rdzc += k; G.rdzf += k; G.zc += k; G.zf += k;
  G.oi += i; G.oj += j; G.ok += k;
}

// volume element at a request's location (Integral(op) = Σ op·V, Oceananigans Reduction(sum!, op * V))
template <class T>
OCB_X double ocb_cell_volume(const DGrid<T>& g, int locz, int k) {
  T dz = (locz == OCB_FACE) ? OCB_LD(g.dzf + k) : OCB_LD(g.dzc + k);
  return (double)g.dx * (double)g.dy * (double)dz;
}
