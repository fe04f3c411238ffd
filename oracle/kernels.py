"""Per-cell kernel functions of the hot path, restated in NumPy (test infrastructure only).

Each function cites the reference file:line it follows (paths relative to
/root/reference).  The Oceananigans pieces they call (closure fluxes, centred
momentum advection, buoyancy projections) are restated from the published
Oceananigans v0.10x source [OCN-mem] in the first half of this file.

Kernel functions have the Julia signature ``f(i, j, k, grid, args...)`` with
``i, j, k`` broadcastable integer index arrays; ``evaluate(f, grid, loc, *args)``
evaluates one over the interior of a field located at ``loc``.
"""
from __future__ import annotations

import numpy as np

from .grid import OField, ConstantField, ZeroField, DiffField  # noqa: F401
from .operators import *  # noqa: F401,F403
from . import operators as _op


# =============================================================================
# Oceananigans restatements [OCN-mem]
# =============================================================================
class Closure:
    """Isotropic scalar-diffusivity closure (ThreeDimensionalFormulation).

    ``nu`` / ``kappa``: a Number (ScalarDiffusivity) or a ccc ``OField``
    (closure_fields.νₑ / κₑ of SmagorinskyLilly / AnisotropicMinimumDissipation).
    ``Pr``: when given with a field ``nu`` and ``kappa is None``, κ = νₑ / Pr
    (Smagorinsky's ``diffusivity(closure, K, id) = K.νₑ / closure.Pr[id]``).
    """

    def __init__(self, nu=0.0, kappa=None, Pr=None):
        self.nu, self.kappa, self.Pr = nu, kappa, Pr


class _DivByPr:
    def __init__(self, f, Pr):
        self.f, self.Pr = f, Pr

    def __getitem__(self, ijk):
        return self.f[ijk] / self.Pr


def _closures(closure):
    return tuple(closure) if isinstance(closure, (tuple, list)) else (closure,)


def _is_num(x):
    return not hasattr(x, "__getitem__")


def nu_at(loc, i, j, k, grid, c: Closure):
    """νᶜᶜᶜ / νᶠᶠᶜ / νᶠᶜᶠ / νᶜᶠᶠ: the constant, or ℑ of the ccc νₑ field to ``loc``."""
    nu = c.nu
    if _is_num(nu):
        return nu
    if loc == "ccc":
        return nu[i, j, k]
    if loc == "ffc":
        return _op.ixy_ffa(i, j, k, grid, nu)
    if loc == "fcf":
        return _op.ixz_faf(i, j, k, grid, nu)
    if loc == "cff":
        return _op.iyz_aff(i, j, k, grid, nu)
    raise ValueError(loc)


def kappa_at(loc, i, j, k, grid, c: Closure):
    kap = c.kappa
    if kap is None:
        kap = _DivByPr(c.nu, c.Pr) if c.Pr is not None else 0.0
    if _is_num(kap):
        return kap
    if loc == "fcc":
        return _op.ix_faa(i, j, k, grid, kap)
    if loc == "cfc":
        return _op.iy_afa(i, j, k, grid, kap)
    if loc == "ccf":
        return _op.iz_aaf(i, j, k, grid, kap)
    raise ValueError(loc)


# strain-rate components Σᵢⱼ at their natural locations (TurbulenceClosures/velocity_tracer_gradients.jl)
def S11(i, j, k, g, u, v, w): return _op.ddx_ccc(i, j, k, g, u)
def S22(i, j, k, g, u, v, w): return _op.ddy_ccc(i, j, k, g, v)
def S33(i, j, k, g, u, v, w): return _op.ddz_ccc(i, j, k, g, w)
def S12(i, j, k, g, u, v, w): return (_op.ddy_ffc(i, j, k, g, u) + _op.ddx_ffc(i, j, k, g, v)) / 2
def S13(i, j, k, g, u, v, w): return (_op.ddz_fcf(i, j, k, g, u) + _op.ddx_fcf(i, j, k, g, w)) / 2
def S23(i, j, k, g, u, v, w): return (_op.ddz_cff(i, j, k, g, v) + _op.ddy_cff(i, j, k, g, w)) / 2


def _visc(loc, sigma):
    def flux(i, j, k, grid, closure, K, clock, fields, b):
        u, v, w = fields["u"], fields["v"], fields["w"]
        tot = None
        for c in _closures(closure):   # closure tuples sum their fluxes
            t = -2 * (nu_at(loc, i, j, k, grid, c) * sigma(i, j, k, grid, u, v, w))
            tot = t if tot is None else tot + t
        return tot
    return flux


viscous_flux_ux = _visc("ccc", S11)
viscous_flux_uy = _visc("ffc", S12)
viscous_flux_uz = _visc("fcf", S13)
viscous_flux_vx = _visc("ffc", S12)
viscous_flux_vy = _visc("ccc", S22)
viscous_flux_vz = _visc("cff", S23)
viscous_flux_wx = _visc("fcf", S13)
viscous_flux_wy = _visc("cff", S23)
viscous_flux_wz = _visc("ccc", S33)


def _diff(loc, deriv):
    def flux(i, j, k, grid, closure, K, idv, c, clock=None, fields=None, b=None):
        tot = None
        for cl in _closures(closure):
            t = -(kappa_at(loc, i, j, k, grid, cl) * deriv(i, j, k, grid, c))
            tot = t if tot is None else tot + t
        return tot
    return flux


diffusive_flux_x = _diff("fcc", _op.ddx_fcc)
diffusive_flux_y = _diff("cfc", _op.ddy_cfc)
diffusive_flux_z = _diff("ccf", _op.ddz_ccf)


def _Ax_q_fcc(i, j, k, g, q, *a): return _op.Ax_fcc(i, j, k, g) * q(i, j, k, g, *a)
def _Ay_q_cfc(i, j, k, g, q, *a): return _op.Ay_cfc(i, j, k, g) * q(i, j, k, g, *a)
def _Az_q_ccf(i, j, k, g, q, *a): return _op.Az_ccf(i, j, k, g) * q(i, j, k, g, *a)


def div_q_c(i, j, k, grid, closure, K, idv, c, *args):
    """∇_dot_qᶜ = 1/V (δx(Ax qˣ) + δy(Ay qʸ) + δz(Az qᶻ))  [OCN-mem TurbulenceClosures]."""
    return 1 / _op.V_ccc(i, j, k, grid) * (
        _op.dx_caa(i, j, k, grid, _Ax_q_fcc, diffusive_flux_x, closure, K, idv, c, *args) +
        _op.dy_aca(i, j, k, grid, _Ay_q_cfc, diffusive_flux_y, closure, K, idv, c, *args) +
        _op.dz_aac(i, j, k, grid, _Az_q_ccf, diffusive_flux_z, closure, K, idv, c, *args))


# --- Centered (2nd-order) momentum advection  [OCN-mem Advection/momentum_advection_operators.jl]
def _Axu(i, j, k, g, u): return _op.Ax_fcc(i, j, k, g) * u[i, j, k]
def _Ayv(i, j, k, g, v): return _op.Ay_cfc(i, j, k, g) * v[i, j, k]
def _Azw(i, j, k, g, w): return _op.Az_ccf(i, j, k, g) * w[i, j, k]


# Centered(order = N) symmetric interpolation [OCN-mem Advection/centered_reconstruction.jl, topologically_conditional_interpolation.jl].
# order 2: the two-point mean.  order 4: (-1, 7, 7, -1)/12 over the four nearest points (the uniform-grid coefficients, which
# `Centered(order = 4)` built without a grid uses on stretched axes too); along a Bounded axis the stencil falls back to order 2
# where it would reach across the wall: a face index m keeps order 4 for H+1 <= m <= N+1-H, a centre index for H <= m <= N+1-H,
# with H = required_halo_size = 2 (outside_symmetric_haloᶠ / ᶜ).
def advection_order(advection):
    if advection is None or advection in ("Centered", "Centered(order=2)", 2):
        return 2
    if advection in ("Centered(order=4)", 4):
        return 4
    raise ValueError(f"advection scheme {advection!r} is not restated")


def sym_interp(i, j, k, g, order, d, target, f, *args):
    """Symmetric interpolation of `f` along axis d (0, 1, 2) to the location `target` ('c': from faces, index m between faces m
    and m+1; 'f': from centres, index m between centres m-1 and m)."""
    idx = (i, j, k)[d]
    sh = lambda s: f(*_op._sh(d, i, j, k, s), g, *args) if callable(f) else f[_op._sh(d, i, j, k, s)]
    lo = 0 if target == "c" else -1                     # the two nearest points: lo, lo + 1
    two = (sh(lo) + sh(lo + 1)) / 2
    if order == 2:
        return two
    four = (7 * (sh(lo) + sh(lo + 1)) - (sh(lo - 1) + sh(lo + 2))) / 12
    if g.topology[d] != "B":
        return four
    N, Hs = g.N[d], 2
    inside = (idx >= (Hs if target == "c" else Hs + 1)) & (idx <= N + 1 - Hs)
    return np.where(inside, four, two)


def _flux_Uu(i, j, k, g, o, U, u): return sym_interp(i, j, k, g, o, 0, "c", _Axu, U[0]) * sym_interp(i, j, k, g, o, 0, "c", u)
def _flux_Vu(i, j, k, g, o, U, u): return sym_interp(i, j, k, g, o, 0, "f", _Ayv, U[1]) * sym_interp(i, j, k, g, o, 1, "f", u)
def _flux_Wu(i, j, k, g, o, U, u): return sym_interp(i, j, k, g, o, 0, "f", _Azw, U[2]) * sym_interp(i, j, k, g, o, 2, "f", u)
def _flux_Uv(i, j, k, g, o, U, v): return sym_interp(i, j, k, g, o, 1, "f", _Axu, U[0]) * sym_interp(i, j, k, g, o, 0, "f", v)
def _flux_Vv(i, j, k, g, o, U, v): return sym_interp(i, j, k, g, o, 1, "c", _Ayv, U[1]) * sym_interp(i, j, k, g, o, 1, "c", v)
def _flux_Wv(i, j, k, g, o, U, v): return sym_interp(i, j, k, g, o, 1, "f", _Azw, U[2]) * sym_interp(i, j, k, g, o, 2, "f", v)
def _flux_Uw(i, j, k, g, o, U, w): return sym_interp(i, j, k, g, o, 2, "f", _Axu, U[0]) * sym_interp(i, j, k, g, o, 0, "f", w)
def _flux_Vw(i, j, k, g, o, U, w): return sym_interp(i, j, k, g, o, 2, "f", _Ayv, U[1]) * sym_interp(i, j, k, g, o, 1, "f", w)
def _flux_Ww(i, j, k, g, o, U, w): return sym_interp(i, j, k, g, o, 2, "c", _Azw, U[2]) * sym_interp(i, j, k, g, o, 2, "c", w)


def div_Uu(i, j, k, g, advection, U, u):
    U = (U["u"], U["v"], U["w"]) if isinstance(U, dict) else U
    o = advection_order(advection)
    return 1 / _op.V_fcc(i, j, k, g) * (_op.dx_faa(i, j, k, g, _flux_Uu, o, U, u) +
                                       _op.dy_aca(i, j, k, g, _flux_Vu, o, U, u) +
                                       _op.dz_aac(i, j, k, g, _flux_Wu, o, U, u))


def div_Uv(i, j, k, g, advection, U, v):
    U = (U["u"], U["v"], U["w"]) if isinstance(U, dict) else U
    o = advection_order(advection)
    return 1 / _op.V_cfc(i, j, k, g) * (_op.dx_caa(i, j, k, g, _flux_Uv, o, U, v) +
                                       _op.dy_afa(i, j, k, g, _flux_Vv, o, U, v) +
                                       _op.dz_aac(i, j, k, g, _flux_Wv, o, U, v))


def div_Uw(i, j, k, g, advection, U, w):
    U = (U["u"], U["v"], U["w"]) if isinstance(U, dict) else U
    o = advection_order(advection)
    return 1 / _op.V_ccf(i, j, k, g) * (_op.dx_caa(i, j, k, g, _flux_Uw, o, U, w) +
                                       _op.dy_aca(i, j, k, g, _flux_Vw, o, U, w) +
                                       _op.dz_aaf(i, j, k, g, _flux_Ww, o, U, w))


# --- buoyancy projections for BuoyancyTracer [OCN-mem BuoyancyFormulations/buoyancy_force.jl]
# ``ghat`` = (ĝ_x, ĝ_y, ĝ_z) = -gravity_unit_vector; (0, 0, 1) for the default NegativeZDirection.
def x_dot_g_b_fcc(i, j, k, g, ghat, b): return ghat[0] * _op.ix_faa(i, j, k, g, b)
def y_dot_g_b_cfc(i, j, k, g, ghat, b): return ghat[1] * _op.iy_afa(i, j, k, g, b)
def z_dot_g_b_ccf(i, j, k, g, ghat, b): return ghat[2] * _op.iz_aaf(i, j, k, g, b)


# =============================================================================
# Oceanostics kernels
# =============================================================================
def psi2(i, j, k, g, psi):                                   # src/KineticEnergyEquation.jl:33
    return psi[i, j, k] ** 2


def fpsi_plus_gphi2(i, j, k, g, f, psi, gg, phi):            # src/KineticEnergyEquation.jl:34
    return (f(i, j, k, g, psi) + gg(i, j, k, g, phi)) ** 2


def fpsi_minus_gphi2(i, j, k, g, f, psi, gg, phi):           # src/FlowDiagnostics.jl:575
    return (f(i, j, k, g, psi) - gg(i, j, k, g, phi)) ** 2


def psif(i, j, k, g, psi, f, *args):                         # src/KineticEnergyEquation.jl:72
    return psi[i, j, k] * f(i, j, k, g, *args)


def kinetic_energy_ccc(i, j, k, g, u, v, w):                 # src/KineticEnergyEquation.jl:37-39
    return (ix_caa(i, j, k, g, psi2, u) + iy_aca(i, j, k, g, psi2, v) + iz_aac(i, j, k, g, psi2, w)) / 2


def ke_advection_ccc(i, j, k, g, velocities, advection=None):   # uᵢ∂ⱼuⱼuᵢᶜᶜᶜ, src/KineticEnergyEquation.jl:147-152 (any Centered order)
    U = velocities
    a = ix_caa(i, j, k, g, psif, U["u"], div_Uu, advection, U, U["u"])
    b = iy_aca(i, j, k, g, psif, U["v"], div_Uv, advection, U, U["v"])
    c = iz_aac(i, j, k, g, psif, U["w"], div_Uw, advection, U, U["w"])
    return a + b + c


def ke_pressure_ccc(i, j, k, g, velocities, p):              # uᵢ∂ᵢpᶜᶜᶜ, src/KineticEnergyEquation.jl:304-309
    a = ix_caa(i, j, k, g, psif, velocities["u"], ddx_fcc, p)
    b = iy_aca(i, j, k, g, psif, velocities["v"], ddy_cfc, p)
    c = iz_aac(i, j, k, g, psif, velocities["w"], ddz_ccf, p)
    return a + b + c


def ke_buoyancy_ccc(i, j, k, g, velocities, ghat, b):        # uᵢbᵢᶜᶜᶜ, src/KineticEnergyEquation.jl:362-367
    a = ix_caa(i, j, k, g, psif, velocities["u"], x_dot_g_b_fcc, ghat, b)
    bb = iy_aca(i, j, k, g, psif, velocities["v"], y_dot_g_b_cfc, ghat, b)
    c = iz_aac(i, j, k, g, psif, velocities["w"], z_dot_g_b_ccf, ghat, b)
    return a + bb + c


# src/KineticEnergyEquation.jl:427-439
def _A_du_t11(i, j, k, g, cl, K, clo, f, b): return -Ax_ccc(i, j, k, g) * dx_caa(i, j, k, g, f["u"]) * viscous_flux_ux(i, j, k, g, cl, K, clo, f, b)
def _A_du_t12(i, j, k, g, cl, K, clo, f, b): return -Ay_ffc(i, j, k, g) * dy_afa(i, j, k, g, f["u"]) * viscous_flux_uy(i, j, k, g, cl, K, clo, f, b)
def _A_du_t13(i, j, k, g, cl, K, clo, f, b): return -Az_fcf(i, j, k, g) * dz_aaf(i, j, k, g, f["u"]) * viscous_flux_uz(i, j, k, g, cl, K, clo, f, b)
def _A_dv_t21(i, j, k, g, cl, K, clo, f, b): return -Ax_ffc(i, j, k, g) * dx_faa(i, j, k, g, f["v"]) * viscous_flux_vx(i, j, k, g, cl, K, clo, f, b)
def _A_dv_t22(i, j, k, g, cl, K, clo, f, b): return -Ay_ccc(i, j, k, g) * dy_aca(i, j, k, g, f["v"]) * viscous_flux_vy(i, j, k, g, cl, K, clo, f, b)
def _A_dv_t23(i, j, k, g, cl, K, clo, f, b): return -Az_cff(i, j, k, g) * dz_aaf(i, j, k, g, f["v"]) * viscous_flux_vz(i, j, k, g, cl, K, clo, f, b)
def _A_dw_t31(i, j, k, g, cl, K, clo, f, b): return -Ax_fcf(i, j, k, g) * dx_faa(i, j, k, g, f["w"]) * viscous_flux_wx(i, j, k, g, cl, K, clo, f, b)
def _A_dw_t32(i, j, k, g, cl, K, clo, f, b): return -Ay_cff(i, j, k, g) * dy_afa(i, j, k, g, f["w"]) * viscous_flux_wy(i, j, k, g, cl, K, clo, f, b)
def _A_dw_t33(i, j, k, g, cl, K, clo, f, b): return -Az_ccc(i, j, k, g) * dz_aac(i, j, k, g, f["w"]) * viscous_flux_wz(i, j, k, g, cl, K, clo, f, b)


def viscous_dissipation_rate_ccc(i, j, k, g, closure_fields, fields, p):   # src/KineticEnergyEquation.jl:441-453
    a = (p["closure"], closure_fields, p.get("clock"), fields, p.get("buoyancy"))
    return (_A_du_t11(i, j, k, g, *a) +
            ixy_cca(i, j, k, g, _A_du_t12, *a) +
            ixz_cac(i, j, k, g, _A_du_t13, *a) +
            ixy_cca(i, j, k, g, _A_dv_t21, *a) +
            _A_dv_t22(i, j, k, g, *a) +
            iyz_acc(i, j, k, g, _A_dv_t23, *a) +
            ixz_cac(i, j, k, g, _A_dw_t31, *a) +
            iyz_acc(i, j, k, g, _A_dw_t32, *a) +
            _A_dw_t33(i, j, k, g, *a)) / V_ccc(i, j, k, g)


# Oceananigans ∂ⱼ_τ₁ⱼ / ∂ⱼ_τ₂ⱼ / ∂ⱼ_τ₃ⱼ [OCN-mem: TurbulenceClosures, viscous flux divergences]: the finite-volume divergence
# of the viscous fluxes at the velocity point, 1/V [δx(Ax τᵢ₁) + δy(Ay τᵢ₂) + δz(Az τᵢ₃)]
def _Ax_t_ccc(i, j, k, g, t, *a): return Ax_ccc(i, j, k, g) * t(i, j, k, g, *a)
def _Ay_t_ffc(i, j, k, g, t, *a): return Ay_ffc(i, j, k, g) * t(i, j, k, g, *a)
def _Az_t_fcf(i, j, k, g, t, *a): return Az_fcf(i, j, k, g) * t(i, j, k, g, *a)
def _Ax_t_ffc(i, j, k, g, t, *a): return Ax_ffc(i, j, k, g) * t(i, j, k, g, *a)
def _Ay_t_ccc(i, j, k, g, t, *a): return Ay_ccc(i, j, k, g) * t(i, j, k, g, *a)
def _Az_t_cff(i, j, k, g, t, *a): return Az_cff(i, j, k, g) * t(i, j, k, g, *a)
def _Ax_t_fcf(i, j, k, g, t, *a): return Ax_fcf(i, j, k, g) * t(i, j, k, g, *a)
def _Ay_t_cff(i, j, k, g, t, *a): return Ay_cff(i, j, k, g) * t(i, j, k, g, *a)
def _Az_t_ccc(i, j, k, g, t, *a): return Az_ccc(i, j, k, g) * t(i, j, k, g, *a)


def div_tau_1(i, j, k, g, *a):     # ∂ⱼ_τ₁ⱼ at fcc
    return (dx_faa(i, j, k, g, _Ax_t_ccc, viscous_flux_ux, *a) + dy_aca(i, j, k, g, _Ay_t_ffc, viscous_flux_uy, *a) +
            dz_aac(i, j, k, g, _Az_t_fcf, viscous_flux_uz, *a)) / _op.V_fcc(i, j, k, g)


def div_tau_2(i, j, k, g, *a):     # ∂ⱼ_τ₂ⱼ at cfc
    return (dx_caa(i, j, k, g, _Ax_t_ffc, viscous_flux_vx, *a) + dy_afa(i, j, k, g, _Ay_t_ccc, viscous_flux_vy, *a) +
            dz_aac(i, j, k, g, _Az_t_cff, viscous_flux_vz, *a)) / _op.V_cfc(i, j, k, g)


def div_tau_3(i, j, k, g, *a):     # ∂ⱼ_τ₃ⱼ at ccf
    return (dx_caa(i, j, k, g, _Ax_t_fcf, viscous_flux_wx, *a) + dy_aca(i, j, k, g, _Ay_t_cff, viscous_flux_wy, *a) +
            dz_aaf(i, j, k, g, _Az_t_ccc, viscous_flux_wz, *a)) / _op.V_ccf(i, j, k, g)


def ke_stress_ccc(i, j, k, g, closure, closure_fields, clock, fields, buoyancy):   # uᵢ∂ⱼ_τᵢⱼᶜᶜᶜ, src/KineticEnergyEquation.jl:191-202
    a = (closure, closure_fields, clock, fields, buoyancy)
    return (_op.ix_caa(i, j, k, g, psif, fields["u"], div_tau_1, *a) + _op.iy_aca(i, j, k, g, psif, fields["v"], div_tau_2, *a) +
            _op.iz_aac(i, j, k, g, psif, fields["w"], div_tau_3, *a))


# --- Coriolis force [OCN-mem Coriolis/constant_cartesian_coriolis.jl; FPlane(f) is f = (0, 0, f): Coriolis/f_plane.jl]
def x_f_cross_U(i, j, k, g, f, U): return f[1] * _op.ixz_fac(i, j, k, g, U["w"]) - f[2] * _op.ixy_fca(i, j, k, g, U["v"])
def y_f_cross_U(i, j, k, g, f, U): return f[2] * _op.ixy_cfa(i, j, k, g, U["u"]) - f[0] * _op.iyz_afc(i, j, k, g, U["w"])
def z_f_cross_U(i, j, k, g, f, U): return f[0] * _op.iyz_acf(i, j, k, g, U["v"]) - f[1] * _op.ixz_caf(i, j, k, g, U["u"])


def _tendency_common(stokes_drift, immersed_bc, background_fields):
    assert stokes_drift is None and immersed_bc is None and background_fields is None, \
        "restated for a model without Stokes drift, immersed boundaries or background fields"


def _hydrostatic_gradient(deriv, i, j, k, g, pHY):
    return 0.0 if pHY is None else deriv(i, j, k, g, pHY)


def u_velocity_tendency(i, j, k, g, advection, coriolis, stokes_drift, closure, u_immersed_bc, buoyancy, background_fields, velocities,
                        tracers, auxiliary_fields, closure_fields, pHY, clock, forcing):
    """Oceananigans NonhydrostaticModels.u_velocity_tendency [OCN-mem velocity_and_tracer_tendencies.jl]:
    -div_𝐯u + x_dot_g_b - x_f_cross_U - ∂x pHY′ - ∂ⱼ_τ₁ⱼ + Fu.  `buoyancy` = (ĝ, name of the buoyancy tracer) or None,
    `coriolis` = (fx, fy, fz) or None."""
    _tendency_common(stokes_drift, u_immersed_bc, background_fields)
    fields = {**velocities, **tracers}
    r = -div_Uu(i, j, k, g, advection, velocities, velocities["u"])
    if buoyancy is not None:
        r = r + x_dot_g_b_fcc(i, j, k, g, buoyancy[0], tracers[buoyancy[1]])
    if coriolis is not None:
        r = r - x_f_cross_U(i, j, k, g, coriolis, velocities)
    r = r - _hydrostatic_gradient(_op.ddx_fcc, i, j, k, g, pHY)
    r = r - div_tau_1(i, j, k, g, closure, closure_fields, clock, fields, buoyancy)
    return r + (forcing(i, j, k, g, clock, fields) if forcing is not None else 0.0)


def v_velocity_tendency(i, j, k, g, advection, coriolis, stokes_drift, closure, v_immersed_bc, buoyancy, background_fields, velocities,
                        tracers, auxiliary_fields, closure_fields, pHY, clock, forcing):
    _tendency_common(stokes_drift, v_immersed_bc, background_fields)
    fields = {**velocities, **tracers}
    r = -div_Uv(i, j, k, g, advection, velocities, velocities["v"])
    if buoyancy is not None:
        r = r + y_dot_g_b_cfc(i, j, k, g, buoyancy[0], tracers[buoyancy[1]])
    if coriolis is not None:
        r = r - y_f_cross_U(i, j, k, g, coriolis, velocities)
    r = r - _hydrostatic_gradient(_op.ddy_cfc, i, j, k, g, pHY)
    r = r - div_tau_2(i, j, k, g, closure, closure_fields, clock, fields, buoyancy)
    return r + (forcing(i, j, k, g, clock, fields) if forcing is not None else 0.0)


def w_velocity_tendency(i, j, k, g, advection, coriolis, stokes_drift, closure, w_immersed_bc, buoyancy, background_fields, velocities,
                        tracers, auxiliary_fields, closure_fields, pHY, clock, forcing):
    """... the vertical buoyancy term only when the model does not separate a hydrostatic pressure anomaly
    (maybe_z_dot_g_bᶜᶜᶠ: zero when pHY′ is a field, since buoyancy then acts through ∇ₕ pHY′)."""
    _tendency_common(stokes_drift, w_immersed_bc, background_fields)
    fields = {**velocities, **tracers}
    r = -div_Uw(i, j, k, g, advection, velocities, velocities["w"])
    if buoyancy is not None and pHY is None:
        r = r + z_dot_g_b_ccf(i, j, k, g, buoyancy[0], tracers[buoyancy[1]])
    if coriolis is not None:
        r = r - z_f_cross_U(i, j, k, g, coriolis, velocities)
    r = r - div_tau_3(i, j, k, g, closure, closure_fields, clock, fields, buoyancy)
    return r + (forcing(i, j, k, g, clock, fields) if forcing is not None else 0.0)


def ke_tendency_ccc(i, j, k, g, advection, coriolis, stokes_drift, closure, u_ibc, v_ibc, w_ibc, buoyancy, background_fields,
                    velocities, tracers, auxiliary_fields, closure_fields, pHY, clock, forcings):   # uᵢGᵢᶜᶜᶜ, src/KineticEnergyEquation.jl:74-95
    common = (buoyancy, background_fields, velocities, tracers, auxiliary_fields, closure_fields, pHY, clock)
    F = forcings or {}
    a = _op.ix_caa(i, j, k, g, psif, velocities["u"], u_velocity_tendency, advection, coriolis, stokes_drift, closure, u_ibc, *common, F.get("u"))
    b = _op.iy_aca(i, j, k, g, psif, velocities["v"], v_velocity_tendency, advection, coriolis, stokes_drift, closure, v_ibc, *common, F.get("v"))
    c = _op.iz_aac(i, j, k, g, psif, velocities["w"], w_velocity_tendency, advection, coriolis, stokes_drift, closure, w_ibc, *common, F.get("w"))
    return a + b + c


# --- Smagorinsky-Lilly eddy viscosity [OCN-mem TurbulenceClosures/.../smagorinsky_lilly.jl: calc_nonlinear_νᶜᶜᶜ, stability]
def _dzb_ccf(i, j, k, g, b): return _op.ddz_ccf(i, j, k, g, b)


def smagorinsky_nu_ccc(i, j, k, g, u, v, w, b, C, Cb=0.0, Pr=1.0):
    """νₑ = ς (C Δᶠ)² √(2 ΣᵢⱼΣᵢⱼ), Δᶠ = ∛(ΔxΔyΔz) at ccc, ς = Σ² == 0 ? 0 : √max(0, 1 - Cb N²⁺ / (Pr Σ²)),
    N²⁺ = max(0, ℑzᵃᵃᶜ ∂zᶜᶜᶠ b).  ΣᵢⱼΣᵢⱼ is the square of strain_rate_tensor_modulus_ccc (src/FlowDiagnostics.jl:431-441).
    Cb = 0 (Oceananigans' `Cb = nothing`): no stratification correction."""
    S2 = strain_rate_tensor_modulus_ccc(i, j, k, g, u, v, w) ** 2
    delta = np.cbrt(_op.Dx_ccc(i, j, k, g) * _op.Dy_ccc(i, j, k, g) * _op.Dz_ccc(i, j, k, g))
    sig = 1.0
    if Cb:
        N2 = np.maximum(0.0, _op.iz_aac(i, j, k, g, _dzb_ccf, b))
        with np.errstate(divide="ignore", invalid="ignore"):
            sig = np.where(S2 == 0, 0.0, np.sqrt(np.maximum(0.0, 1.0 - Cb * N2 / (Pr * S2))))
    return sig * (C * delta) ** 2 * np.sqrt(2.0 * S2)


# --- Centered (2nd-order) tracer advection  [OCN-mem Advection/tracer_advection_operators.jl]: advective fluxes
# Ax u ℑx c (fcc), Ay v ℑy c (cfc), Az w ℑz c (ccf) and their divergence at the cell centre
def _adv_flux_x(i, j, k, g, o, U, c): return _op.Ax_fcc(i, j, k, g) * U[0][i, j, k] * sym_interp(i, j, k, g, o, 0, "f", c)
def _adv_flux_y(i, j, k, g, o, U, c): return _op.Ay_cfc(i, j, k, g) * U[1][i, j, k] * sym_interp(i, j, k, g, o, 1, "f", c)
def _adv_flux_z(i, j, k, g, o, U, c): return _op.Az_ccf(i, j, k, g) * U[2][i, j, k] * sym_interp(i, j, k, g, o, 2, "f", c)


def div_Uc(i, j, k, g, advection, U, c):
    o = advection_order(advection)
    return (_op.dx_caa(i, j, k, g, _adv_flux_x, o, U, c) + _op.dy_aca(i, j, k, g, _adv_flux_y, o, U, c) +
            _op.dz_aac(i, j, k, g, _adv_flux_z, o, U, c)) / _op.V_ccc(i, j, k, g)


def tracer_tendency(i, j, k, g, idx, name, advection, closure, c_immersed_bc, buoyancy, biogeochemistry, background_fields,
                    velocities, tracers, auxiliary_fields, closure_fields, clock, forcing):
    """Oceananigans NonhydrostaticModels.tracer_tendency [OCN-mem] for a model without background fields, biogeochemistry or
    immersed boundaries (those arguments must be None here): -div_Uc - ∇_dot_qᶜ + forcing."""
    assert c_immersed_bc is None and biogeochemistry is None and background_fields is None
    c = tracers[name]
    fields = {**velocities, **tracers}
    U = (velocities["u"], velocities["v"], velocities["w"])
    F = forcing(i, j, k, g, clock, fields) if forcing is not None else 0.0
    return -div_Uc(i, j, k, g, advection, U, c) - div_q_c(i, j, k, g, closure, closure_fields, idx, c, clock, fields, buoyancy) + F


def c_dt_c_ccc(i, j, k, g, idx, name, *args):                  # c∂ₜcᶜᶜᶜ, src/TracerVarianceEquation.jl:18-39
    tracers = args[7]                                          # (advection, closure, c_immersed_bc, buoyancy, biogeochemistry, background_fields, velocities, tracers, ...)
    return 2 * tracers[name][i, j, k] * tracer_tendency(i, j, k, g, idx, name, *args)


class ContinuousForcing:
    """Oceananigans ContinuousForcing [OCN-mem]: F(i, j, k, grid, clock, fields) = func(x, y, z, t, deps..., [parameters]) with the
    nodes and the dependency fields taken at the forced field's own location `loc` (dependencies located elsewhere would be
    interpolated there by Oceananigans; the cases restated here depend on the forced field itself)."""

    def __init__(self, func, loc, field_dependencies=(), parameters=None):
        self.func, self.loc, self.deps, self.parameters = func, loc, tuple(field_dependencies), parameters

    def __call__(self, i, j, k, g, clock, fields):
        x, y, z = g.node(0, self.loc[0], i), g.node(1, self.loc[1], j), g.node(2, self.loc[2], k)
        args = [x, y, z, getattr(clock, "time", 0.0) if clock is not None else 0.0] + [fields[n][i, j, k] for n in self.deps]
        if self.parameters is not None:
            args.append(self.parameters)
        return self.func(*args)


def _zero_forcing(i, j, k, g, clock, fields):
    return 0.0


def ke_forcing_ccc(i, j, k, g, forcings, clock, fields):       # uᵢFᵤᵢᶜᶜᶜ, src/KineticEnergyEquation.jl:251-260
    Fu, Fv, Fw = (forcings.get(n) or _zero_forcing for n in ("u", "v", "w"))
    return (_op.ix_caa(i, j, k, g, psif, fields["u"], Fu, clock, fields) + _op.iy_aca(i, j, k, g, psif, fields["v"], Fv, clock, fields) +
            _op.iz_aac(i, j, k, g, psif, fields["w"], Fw, clock, fields))


def nu_ccc_total(i, j, k, g, closure):                       # _νᶜᶜᶜ, src/Oceanostics.jl:201-208
    tot = None
    for c in _closures(closure):
        t = nu_at("ccc", i, j, k, g, c)
        tot = t if tot is None else tot + t
    return tot


def isotropic_viscous_dissipation_rate_ccc(i, j, k, g, u, v, w, p):   # src/KineticEnergyEquation.jl:497-510
    sxx = ddx_ccc(i, j, k, g, u) ** 2
    syy = ddy_ccc(i, j, k, g, v) ** 2
    szz = ddz_ccc(i, j, k, g, w) ** 2
    sxy = ixy_cca(i, j, k, g, fpsi_plus_gphi2, ddy_ffc, u, ddx_ffc, v) / 4
    sxz = ixz_cac(i, j, k, g, fpsi_plus_gphi2, ddz_fcf, u, ddx_fcf, w) / 4
    syz = iyz_acc(i, j, k, g, fpsi_plus_gphi2, ddz_cff, v, ddy_cff, w) / 4
    nu = nu_ccc_total(i, j, k, g, p["closure"])
    return 2 * nu * (sxx + syy + szz + 2 * (sxy + sxz + syz))


# --- TurbulentKineticEnergyEquation --------------------------------------------------------------
def psip2(i, j, k, g, psi, psibar):                          # ψ′², src/TurbulentKineticEnergyEquation.jl:24-25
    return (psi[i, j, k] - psibar[i, j, k]) ** 2


def turbulent_kinetic_energy_ccc(i, j, k, g, u, v, w, U, V, W):   # :28-30
    return (ix_caa(i, j, k, g, psip2, u, U) + iy_aca(i, j, k, g, psip2, v, V) + iz_aac(i, j, k, g, psip2, w, W)) / 2


def shear_production_rate_x_ccc(i, j, k, g, up, vp, wp, U, V, W):   # :101-117
    up_int = ix_caa(i, j, k, g, up)
    dxU = ddx_ccc(i, j, k, g, U)
    upup = ix_caa(i, j, k, g, psi2, up)
    dxV = ixy_cca(i, j, k, g, ddx_ffc, V)
    vpup = iy_aca(i, j, k, g, vp) * up_int
    dxW = ixz_cac(i, j, k, g, ddx_fcf, W)
    wpup = iz_aac(i, j, k, g, wp) * up_int
    return -(upup * dxU + vpup * dxV + wpup * dxW)


def shear_production_rate_y_ccc(i, j, k, g, up, vp, wp, U, V, W):   # :163-179
    vp_int = iy_aca(i, j, k, g, vp)
    dyU = ixy_cca(i, j, k, g, ddy_ffc, U)
    upvp = ix_caa(i, j, k, g, up) * vp_int
    dyV = ddy_ccc(i, j, k, g, V)
    vpvp = iy_aca(i, j, k, g, psi2, vp)
    dyW = iyz_acc(i, j, k, g, ddy_cff, W)
    wpvp = iz_aac(i, j, k, g, wp) * vp_int
    return -(upvp * dyU + vpvp * dyV + wpvp * dyW)


def shear_production_rate_z_ccc(i, j, k, g, up, vp, wp, U, V, W):   # :224-240
    wp_int = iz_aac(i, j, k, g, wp)
    dzU = ixz_cac(i, j, k, g, ddz_fcf, U)
    upwp = ix_caa(i, j, k, g, up) * wp_int
    dzV = iyz_acc(i, j, k, g, ddz_cff, V)
    vpwp = iy_aca(i, j, k, g, vp) * wp_int
    dzW = ddz_ccc(i, j, k, g, W)
    wpwp = iz_aac(i, j, k, g, psi2, wp)
    return -(upwp * dzU + vpwp * dzV + wpwp * dzW)


def shear_production_rate_ccc(*args):                        # :285-289
    return (shear_production_rate_x_ccc(*args) + shear_production_rate_y_ccc(*args) +
            shear_production_rate_z_ccc(*args))


# --- TracerVarianceEquation ------------------------------------------------------------------------
def c_div_q_c(i, j, k, g, closure, K, idv, c, *args):        # c∇_dot_qᶜ, src/TracerVarianceEquation.jl:93-98
    return 2 * c[i, j, k] * div_q_c(i, j, k, g, closure, K, idv, c, *args)


def _Ax_dc_q1(i, j, k, g, cl, K, idv, c, *a): return -Ax_fcc(i, j, k, g) * dx_faa(i, j, k, g, c) * diffusive_flux_x(i, j, k, g, cl, K, idv, c, *a)   # :146-147
def _Ay_dc_q2(i, j, k, g, cl, K, idv, c, *a): return -Ay_cfc(i, j, k, g) * dy_afa(i, j, k, g, c) * diffusive_flux_y(i, j, k, g, cl, K, idv, c, *a)   # :150-151
def _Az_dc_q3(i, j, k, g, cl, K, idv, c, *a): return -Az_ccf(i, j, k, g) * dz_aaf(i, j, k, g, c) * diffusive_flux_z(i, j, k, g, cl, K, idv, c, *a)   # :154-155


def tracer_variance_dissipation_rate_ccc(i, j, k, g, *args):   # :157-161
    return 2 * (ix_caa(i, j, k, g, _Ax_dc_q1, *args) +
                iy_aca(i, j, k, g, _Ay_dc_q2, *args) +
                iz_aac(i, j, k, g, _Az_dc_q3, *args)) / V_ccc(i, j, k, g)


# --- AvailablePotentialEnergyEquation ---------------------------------------------------------------
def local_ape_ccc(i, j, k, g, psi_delta, b, zstar):          # src/AvailablePotentialEnergyEquation.jl:42
    return psi_delta[i, j, k] - b[i, j, k] * (g.node(2, "c", k) - zstar[i, j, k])


def upsilon_ccc(i, j, k, g, zstar):                          # src/AvailablePotentialEnergyEquation.jl:126
    return zstar[i, j, k] - g.node(2, "c", k)


def _Ax_dY_q1(i, j, k, g, Y, cl, K, idv, c, *a): return -Ax_fcc(i, j, k, g) * dx_faa(i, j, k, g, Y) * diffusive_flux_x(i, j, k, g, cl, K, idv, c, *a)   # :202-203
def _Ay_dY_q2(i, j, k, g, Y, cl, K, idv, c, *a): return -Ay_cfc(i, j, k, g) * dy_afa(i, j, k, g, Y) * diffusive_flux_y(i, j, k, g, cl, K, idv, c, *a)   # :205-206
def _Az_dY_q3(i, j, k, g, Y, cl, K, idv, c, *a): return -Az_ccf(i, j, k, g) * dz_aaf(i, j, k, g, Y) * diffusive_flux_z(i, j, k, g, cl, K, idv, c, *a)   # :208-209


def ape_dissipation_rate_ccc(i, j, k, g, *args):             # src/AvailablePotentialEnergyEquation.jl:211-215
    return (ix_caa(i, j, k, g, _Ax_dY_q1, *args) +
            iy_aca(i, j, k, g, _Ay_dY_q2, *args) +
            iz_aac(i, j, k, g, _Az_dY_q3, *args)) / V_ccc(i, j, k, g)


# --- FlowDiagnostics --------------------------------------------------------------------------------
def w2_from_u_tilted_ccc(i, j, k, g, u, v, w, vdir):          # src/FlowDiagnostics.jl:36-41
    uh = ix_caa(i, j, k, g, u)
    vh = iy_aca(i, j, k, g, v)
    wh = iz_aac(i, j, k, g, w)
    return (uh * vdir[0] + vh * vdir[1] + wh * vdir[2]) ** 2


def uh_norm_ccc(i, j, k, g, u, v, w, vdir):                   # :48-53
    u2 = ix_caa(i, j, k, g, psi2, u)
    v2 = iy_aca(i, j, k, g, psi2, v)
    w2 = iz_aac(i, j, k, g, psi2, w)
    return np.sqrt(np.maximum(0.0, u2 + v2 + w2 - w2_from_u_tilted_ccc(i, j, k, g, u, v, w, vdir)))


def richardson_number_ccf(i, j, k, g, u, v, w, b, vdir):     # :55-68
    dbdx = ixz_caf(i, j, k, g, ddx_fcc, b)
    dbdy = iyz_acf(i, j, k, g, ddy_cfc, b)
    dbdz_ = ddz_ccf(i, j, k, g, b)
    dbdz = dbdx * vdir[0] + dbdy * vdir[1] + dbdz_ * vdir[2]
    dudx = ix_caa(i, j, k, g, ddx_fcc, uh_norm_ccc, u, v, w, vdir)
    dudy = iy_aca(i, j, k, g, ddy_cfc, uh_norm_ccc, u, v, w, vdir)
    dudz_ = ddz_ccf(i, j, k, g, uh_norm_ccc, u, v, w, vdir)
    dudz = dudx * vdir[0] + dudy * vdir[1] + dudz_ * vdir[2]
    with np.errstate(divide="ignore", invalid="ignore"):
        return dbdz / dudz ** 2


def rossby_number_fff(i, j, k, g, u, v, w, p):               # :124-138
    dwdy = ix_faa(i, j, k, g, ddy_cff, w)
    dvdz = ix_faa(i, j, k, g, ddz_cff, v)
    wx = (dwdy + p["dWdy_bg"]) - (dvdz + p["dVdz_bg"])
    dudz = iy_afa(i, j, k, g, ddz_fcf, u)
    dwdx = iy_afa(i, j, k, g, ddx_fcf, w)
    wy = (dudz + p["dUdz_bg"]) - (dwdx + p["dWdx_bg"])
    dvdx = iz_aaf(i, j, k, g, ddx_ffc, v)
    dudy = iz_aaf(i, j, k, g, ddy_ffc, u)
    wz = (dvdx + p["dVdx_bg"]) - (dudy + p["dUdy_bg"])
    return (wx * p["fx"] + wy * p["fy"] + wz * p["fz"]) / (p["fx"] ** 2 + p["fy"] ** 2 + p["fz"] ** 2)


def potential_vorticity_in_thermal_wind_fff(i, j, k, g, u, v, b, f):   # :200-214
    dVdx = iz_aaf(i, j, k, g, ddx_ffc, v)
    dUdy = iz_aaf(i, j, k, g, ddy_ffc, u)
    dbdz = ixy_ffa(i, j, k, g, ddz_ccf, b)
    pv_barot = (f + dVdx - dUdy) * dbdz
    dUdz = iy_afa(i, j, k, g, ddz_fcf, u)
    dVdz = ix_faa(i, j, k, g, ddz_cff, v)
    pv_baroc = -f * (dUdz ** 2 + dVdz ** 2)
    return pv_barot + pv_baroc


def ertel_potential_vorticity_fff(i, j, k, g, u, v, w, b, fx, fy, fz):   # :216-233
    dWdy = ix_faa(i, j, k, g, ddy_cff, w)
    dVdz = ix_faa(i, j, k, g, ddz_cff, v)
    dbdx = iyz_aff(i, j, k, g, ddx_fcc, b)
    pv_x = (fx + dWdy - dVdz) * dbdx
    dUdz = iy_afa(i, j, k, g, ddz_fcf, u)
    dWdx = iy_afa(i, j, k, g, ddx_fcf, w)
    dbdy = ixz_faf(i, j, k, g, ddy_cfc, b)
    pv_y = (fy + dUdz - dWdx) * dbdy
    dVdx = iz_aaf(i, j, k, g, ddx_ffc, v)
    dUdy = iz_aaf(i, j, k, g, ddy_ffc, u)
    dbdz = ixy_ffa(i, j, k, g, ddz_ccf, b)
    pv_z = (fz + dVdx - dUdy) * dbdz
    return pv_x + pv_y + pv_z


def directional_ertel_potential_vorticity_fff(i, j, k, g, u, v, w, b, p):   # :358-380
    dWdy = ix_faa(i, j, k, g, ddy_cff, w)
    dVdz = ix_faa(i, j, k, g, ddz_cff, v)
    ox = dWdy - dVdz
    dUdz = iy_afa(i, j, k, g, ddz_fcf, u)
    dWdx = iy_afa(i, j, k, g, ddx_fcf, w)
    oy = dUdz - dWdx
    dVdx = iz_aaf(i, j, k, g, ddx_ffc, v)
    dUdy = iz_aaf(i, j, k, g, ddy_ffc, u)
    oz = dVdx - dUdy
    dbdx = iyz_aff(i, j, k, g, ddx_fcc, b)
    dbdy = ixz_faf(i, j, k, g, ddy_cfc, b)
    dbdz = ixy_ffa(i, j, k, g, ddz_ccf, b)
    o_dir = ox * p["dir_x"] + oy * p["dir_y"] + oz * p["dir_z"]
    db_dir = dbdx * p["dir_x"] + dbdy * p["dir_y"] + dbdz * p["dir_z"]
    return (p["f_dir"] + o_dir) * db_dir


def strain_rate_tensor_modulus_ccc(i, j, k, g, u, v, w):      # :431-441
    sxx = ddx_ccc(i, j, k, g, u) ** 2
    syy = ddy_ccc(i, j, k, g, v) ** 2
    szz = ddz_ccc(i, j, k, g, w) ** 2
    sxy = ixy_cca(i, j, k, g, fpsi_plus_gphi2, ddy_ffc, u, ddx_ffc, v) / 4
    sxz = ixz_cac(i, j, k, g, fpsi_plus_gphi2, ddz_fcf, u, ddx_fcf, w) / 4
    syz = iyz_acc(i, j, k, g, fpsi_plus_gphi2, ddz_cff, v, ddy_cff, w) / 4
    return np.sqrt(sxx + syy + szz + 2 * (sxy + sxz + syz))


def strain_rate_tensor_component(I, J):                       # StrainRateTensorKernel{I,J}, :484-491
    table = {(1, 1): lambda i, j, k, g, u: ddx_ccc(i, j, k, g, u),
             (2, 2): lambda i, j, k, g, v: ddy_ccc(i, j, k, g, v),
             (3, 3): lambda i, j, k, g, w: ddz_ccc(i, j, k, g, w),
             (1, 2): lambda i, j, k, g, u, v: (ddy_ffc(i, j, k, g, u) + ddx_ffc(i, j, k, g, v)) / 2,
             (1, 3): lambda i, j, k, g, u, w: (ddz_fcf(i, j, k, g, u) + ddx_fcf(i, j, k, g, w)) / 2,
             (2, 3): lambda i, j, k, g, v, w: (ddz_cff(i, j, k, g, v) + ddy_cff(i, j, k, g, w)) / 2}
    return table[(I, J)]


def vorticity_tensor_modulus_ccc(i, j, k, g, u, v, w):        # :577-587
    oxy = ixy_cca(i, j, k, g, fpsi_minus_gphi2, ddy_ffc, u, ddx_ffc, v) / 4
    oxz = ixz_cac(i, j, k, g, fpsi_minus_gphi2, ddz_fcf, u, ddx_fcf, w) / 4
    oyz = iyz_acc(i, j, k, g, fpsi_minus_gphi2, ddz_cff, v, ddy_cff, w) / 4
    oyx = ixy_cca(i, j, k, g, fpsi_minus_gphi2, ddx_ffc, v, ddy_ffc, u) / 4
    ozx = ixz_cac(i, j, k, g, fpsi_minus_gphi2, ddx_fcf, w, ddz_fcf, u) / 4
    ozy = iyz_acc(i, j, k, g, fpsi_minus_gphi2, ddy_cff, w, ddz_cff, v) / 4
    return np.sqrt(oxy + oxz + oyz + oyx + ozx + ozy)


def vorticity_tensor_component(I, J):                         # VorticityTensorKernel{I,J}, :630-634
    table = {(1, 2): lambda i, j, k, g, u, v: (ddy_ffc(i, j, k, g, u) - ddx_ffc(i, j, k, g, v)) / 2,
             (1, 3): lambda i, j, k, g, u, w: (ddz_fcf(i, j, k, g, u) - ddx_fcf(i, j, k, g, w)) / 2,
             (2, 3): lambda i, j, k, g, v, w: (ddz_cff(i, j, k, g, v) - ddy_cff(i, j, k, g, w)) / 2}
    return table[(I, J)]


def Q_velocity_gradient_tensor_invariant_ccc(i, j, k, g, u, v, w):   # :708-712
    S2 = strain_rate_tensor_modulus_ccc(i, j, k, g, u, v, w) ** 2
    O2 = vorticity_tensor_modulus_ccc(i, j, k, g, u, v, w) ** 2
    return (O2 - S2) / 2


def stress_tensor_component(I, J, collocated=False):          # StressTensorKernel{I,J,Collocated}, :723-733
    if I == J and collocated:
        return {1: lambda i, j, k, g, u: ix_caa(i, j, k, g, u) ** 2,
                2: lambda i, j, k, g, v: iy_aca(i, j, k, g, v) ** 2,
                3: lambda i, j, k, g, w: iz_aac(i, j, k, g, w) ** 2}[I]
    if I == J:
        return lambda i, j, k, g, q: q[i, j, k] ** 2
    table = {(1, 2): lambda i, j, k, g, u, v: iy_afa(i, j, k, g, u) * ix_faa(i, j, k, g, v),
             (1, 3): lambda i, j, k, g, u, w: iz_aaf(i, j, k, g, u) * ix_faa(i, j, k, g, w),
             (2, 3): lambda i, j, k, g, v, w: iz_aaf(i, j, k, g, v) * iy_afa(i, j, k, g, w)}
    return table[(I, J)]


# --- FilteredKineticEnergyEquation / SubFilterKineticEnergyEquation ----------------------------------
# src/FilteredKineticEnergyEquation.jl:275-283
def _du_t11(i, j, k, g, ub, t): return -Ax_ccc(i, j, k, g) * dx_caa(i, j, k, g, ub) * t[i, j, k]
def _du_t12(i, j, k, g, ub, t): return -Ay_ffc(i, j, k, g) * dy_afa(i, j, k, g, ub) * t[i, j, k]
def _du_t13(i, j, k, g, ub, t): return -Az_fcf(i, j, k, g) * dz_aaf(i, j, k, g, ub) * t[i, j, k]
def _dv_t21(i, j, k, g, vb, t): return -Ax_ffc(i, j, k, g) * dx_faa(i, j, k, g, vb) * t[i, j, k]
def _dv_t22(i, j, k, g, vb, t): return -Ay_ccc(i, j, k, g) * dy_aca(i, j, k, g, vb) * t[i, j, k]
def _dv_t23(i, j, k, g, vb, t): return -Az_cff(i, j, k, g) * dz_aaf(i, j, k, g, vb) * t[i, j, k]
def _dw_t31(i, j, k, g, wb, t): return -Ax_fcf(i, j, k, g) * dx_faa(i, j, k, g, wb) * t[i, j, k]
def _dw_t32(i, j, k, g, wb, t): return -Ay_cff(i, j, k, g) * dy_afa(i, j, k, g, wb) * t[i, j, k]
def _dw_t33(i, j, k, g, wb, t): return -Az_ccc(i, j, k, g) * dz_aac(i, j, k, g, wb) * t[i, j, k]


def filtered_dissipation_rate_ccc(i, j, k, g, fv, ff):       # :288-299
    return (_du_t11(i, j, k, g, fv["u"], ff["t11"]) +
            ixy_cca(i, j, k, g, _du_t12, fv["u"], ff["t12"]) +
            ixz_cac(i, j, k, g, _du_t13, fv["u"], ff["t13"]) +
            ixy_cca(i, j, k, g, _dv_t21, fv["v"], ff["t21"]) +
            _dv_t22(i, j, k, g, fv["v"], ff["t22"]) +
            iyz_acc(i, j, k, g, _dv_t23, fv["v"], ff["t23"]) +
            ixz_cac(i, j, k, g, _dw_t31, fv["w"], ff["t31"]) +
            iyz_acc(i, j, k, g, _dw_t32, fv["w"], ff["t32"]) +
            _dw_t33(i, j, k, g, fv["w"], ff["t33"])) / V_ccc(i, j, k, g)


def subfilter_kinetic_energy_ccc(i, j, k, g, kbar, ub, vb, wb):   # src/SubFilterKineticEnergyEquation.jl:29
    return kbar[i, j, k] - kinetic_energy_ccc(i, j, k, g, ub, vb, wb)


def minus_bzstar_ccc(i, j, k, g, b, zstar):                  # src/BackgroundPotentialEnergyEquation.jl:852
    return -b[i, j, k] * zstar[i, j, k]


# `@at (Center, Center, Center)` interpolation of a field from its own location (Oceananigans
# AbstractOperations/at.jl: one two-point ℑ per axis whose location differs)
def to_center(i, j, k, g, f, loc):
    def ev(i, j, k, g):
        return f[i, j, k]
    fn = ev
    if loc[0] == "f":
        fn = (lambda inner: lambda i, j, k, g: ix_caa(i, j, k, g, inner))(fn)
    if loc[1] == "f":
        fn = (lambda inner: lambda i, j, k, g: iy_aca(i, j, k, g, inner))(fn)
    if loc[2] == "f":
        fn = (lambda inner: lambda i, j, k, g: iz_aac(i, j, k, g, inner))(fn)
    return fn(i, j, k, g)


# =============================================================================
def evaluate(f, grid, loc, *args):
    """Evaluate kernel function ``f`` over the interior of a field located at ``loc``."""
    i, j, k = grid.indices(loc)
    out = f(i, j, k, grid, *args)
    return np.broadcast_to(out, grid.extent_of(loc)).astype(grid.dtype, copy=True)


def integral(values, grid, loc):
    """``Integral(op)`` = Σ op·V over the interior (Oceananigans Reduction(sum!, op * V))."""
    i, j, k = grid.indices(loc)
    return float(np.sum(values.astype(np.float64) * grid.V(loc, i, j, k)))


def average(values, grid, loc):
    i, j, k = grid.indices(loc)
    V = np.broadcast_to(grid.V(loc, i, j, k), values.shape)
    return float(np.sum(values.astype(np.float64) * V) / np.sum(V.astype(np.float64)))
