"""Pin the CPU oracle against every known-answer test / golden vector the reference's own test-suite
and doctests hold for the hot path.  CPU only.

Each test names the reference test it restates (paths relative to /root/reference).
"""
import numpy as np
import pytest

from oracle.grid import OGrid, OField, ConstantField, ZeroField, DiffField, average_dims, materialize
from oracle import kernels as K
from oracle import filters as OF
from oracle import bpe as OB

N = 6


def regular_grid(dtype=np.float64):
    return OGrid((N, N, N), extent=(1, 1, 1), dtype=dtype)          # test/test_utils.jl:12-15


def stretched_grid(dtype=np.float64):                               # test/test_utils.jl:17-24
    S = 0.99
    f_asin = lambda k: -np.arcsin(S * (2 * k - N - 2) / N) / np.pi + 0.5
    F1, F2 = f_asin(1), f_asin(N + 1)
    z_faces = lambda k: ((F1 + F2) / 2 - f_asin(k)) / (F1 - F2)
    return OGrid((N, N, N), x=(0, 1), y=(0, 1), z=z_faces, dtype=dtype)


GRIDS = {"regular": regular_grid, "stretched": stretched_grid}


def velocities(g, u0=0.0, v0=0.0, w0=0.0):
    u = OField(g, "fcc").set(u0).fill_halo()
    v = OField(g, "cfc").set(v0).fill_halo()
    w = OField(g, "ccf").set(w0).fill_halo()
    return u, v, w


def make_closures(g, rng):
    nu_e = OField(g, "ccc")
    nu_e.interior[...] = 1e-3 * (1.0 + rng.random(nu_e.interior.shape))
    nu_e.fill_halo()
    return {"scalar": K.Closure(nu=1e-6, kappa=1e-7),
            "smagorinsky_like": K.Closure(nu=nu_e, Pr=1.0),
            "tuple": (K.Closure(nu=1e-6, kappa=1e-7), K.Closure(nu=nu_e, kappa=nu_e))}


def nu_ccc(g, closure, idx):
    i, j, k = (np.array([[[idx[0] + 1]]]), np.array([[[idx[1] + 1]]]), np.array([[[idx[2] + 1]]]))
    return float(np.ravel(K.nu_ccc_total(i, j, k, g, closure))[0])


IDX = (N // 2 - 1, N // 2 - 1, N // 2 - 1)   # (Nx÷2, Ny÷2, Nz÷2), 0-based


# ----------------------------------------------------------------------------- canonical flows
@pytest.mark.parametrize("gname", GRIDS)
@pytest.mark.parametrize("cname", ["scalar", "smagorinsky_like", "tuple"])
def test_uniform_strain_flow(gname, cname):
    """test/test_canonical_flows.jl:44-93."""
    g = GRIDS[gname]()
    cl = make_closures(g, np.random.default_rng(0))[cname]
    a = 3.0
    u, v, w = velocities(g, lambda x, y, z: a * x, lambda x, y, z: -a * y)
    S = K.evaluate(K.strain_rate_tensor_modulus_ccc, g, "ccc", u, v, w)
    Om = K.evaluate(K.vorticity_tensor_modulus_ccc, g, "ccc", u, v, w)
    q = K.evaluate(K.Q_velocity_gradient_tensor_invariant_ccc, g, "ccc", u, v, w)
    eps = K.evaluate(K.viscous_dissipation_rate_ccc, g, "ccc", None, dict(u=u, v=v, w=w), dict(closure=cl))
    assert S[IDX] == pytest.approx(np.sqrt(2) * a)
    assert abs(Om[IDX]) <= 10 * np.finfo(float).eps
    assert q[IDX] == pytest.approx((Om[IDX] ** 2 - S[IDX] ** 2) / 2)
    assert q[IDX] == pytest.approx(-a ** 2)
    assert eps[IDX] == pytest.approx(2 * nu_ccc(g, cl, IDX) * S[IDX] ** 2)
    comps = {(1, 1): ("ccc", (u,)), (2, 2): ("ccc", (v,)), (3, 3): ("ccc", (w,)),
             (1, 2): ("ffc", (u, v)), (1, 3): ("fcf", (u, w)), (2, 3): ("cff", (v, w))}
    vals = {ij: K.evaluate(K.strain_rate_tensor_component(*ij), g, loc, *args)[IDX] for ij, (loc, args) in comps.items()}
    assert vals[(1, 1)] == pytest.approx(a) and vals[(2, 2)] == pytest.approx(-a)
    for ij in [(3, 3), (1, 2), (1, 3), (2, 3)]:
        assert abs(vals[ij]) <= 10 * np.finfo(float).eps
    SS = vals[(1, 1)] ** 2 + vals[(2, 2)] ** 2 + vals[(3, 3)] ** 2 + 2 * (vals[(1, 2)] ** 2 + vals[(1, 3)] ** 2 + vals[(2, 3)] ** 2)
    assert np.sqrt(SS) == pytest.approx(S[IDX])


@pytest.mark.parametrize("gname", GRIDS)
def test_solid_body_rotation_flow(gname):
    """test/test_canonical_flows.jl:99-147."""
    g = GRIDS[gname]()
    zeta = 3.0
    u, v, w = velocities(g, lambda x, y, z: zeta * y / 2, lambda x, y, z: -zeta * x / 2)
    S = K.evaluate(K.strain_rate_tensor_modulus_ccc, g, "ccc", u, v, w)
    Om = K.evaluate(K.vorticity_tensor_modulus_ccc, g, "ccc", u, v, w)
    q = K.evaluate(K.Q_velocity_gradient_tensor_invariant_ccc, g, "ccc", u, v, w)
    eps = K.evaluate(K.viscous_dissipation_rate_ccc, g, "ccc", None, dict(u=u, v=v, w=w), dict(closure=K.Closure(nu=1.0)))
    assert abs(S[IDX]) <= 1e-12
    assert Om[IDX] == pytest.approx(zeta / np.sqrt(2))
    assert q[IDX] == pytest.approx(zeta ** 2 / 4)
    assert eps[IDX] == pytest.approx(0, abs=1e-12)
    o12 = K.evaluate(K.vorticity_tensor_component(1, 2), g, "ffc", u, v)[IDX]
    o13 = K.evaluate(K.vorticity_tensor_component(1, 3), g, "fcf", u, w)[IDX]
    o23 = K.evaluate(K.vorticity_tensor_component(2, 3), g, "cff", v, w)[IDX]
    assert o12 == pytest.approx(zeta / 2) and abs(o13) < 1e-15 and abs(o23) < 1e-15
    assert np.sqrt(2 * (o12 ** 2 + o13 ** 2 + o23 ** 2)) == pytest.approx(Om[IDX])


@pytest.mark.parametrize("gname", GRIDS)
def test_uniform_shear_flow(gname):
    """test/test_canonical_flows.jl:153-186."""
    g = GRIDS[gname]()
    sig = 3.0
    u, v, w = velocities(g, lambda x, y, z: sig * y)
    S = K.evaluate(K.strain_rate_tensor_modulus_ccc, g, "ccc", u, v, w)
    Om = K.evaluate(K.vorticity_tensor_modulus_ccc, g, "ccc", u, v, w)
    q = K.evaluate(K.Q_velocity_gradient_tensor_invariant_ccc, g, "ccc", u, v, w)
    eps = K.evaluate(K.viscous_dissipation_rate_ccc, g, "ccc", None, dict(u=u, v=v, w=w), dict(closure=K.Closure(nu=1.0)))
    assert S[IDX] == pytest.approx(sig / np.sqrt(2)) and Om[IDX] == pytest.approx(sig / np.sqrt(2))
    assert abs(q[IDX]) <= 4 * np.finfo(float).eps * sig ** 2
    assert eps[IDX] == pytest.approx(2 * 1.0 * S[IDX] ** 2)


def test_uniform_flow_stress_tensor():
    """test/test_canonical_flows.jl:192-222."""
    g = regular_grid()
    U, V = 2.0, -3.0
    u, v, w = velocities(g, U, V)
    t = {}
    t[11] = K.evaluate(K.stress_tensor_component(1, 1), g, "fcc", u)[IDX]
    t[22] = K.evaluate(K.stress_tensor_component(2, 2), g, "cfc", v)[IDX]
    t[33] = K.evaluate(K.stress_tensor_component(3, 3), g, "ccf", w)[IDX]
    t[12] = K.evaluate(K.stress_tensor_component(1, 2), g, "ffc", u, v)[IDX]
    t[13] = K.evaluate(K.stress_tensor_component(1, 3), g, "fcf", u, w)[IDX]
    t[23] = K.evaluate(K.stress_tensor_component(2, 3), g, "cff", v, w)[IDX]
    assert t[11] == pytest.approx(U ** 2) and t[22] == pytest.approx(V ** 2) and t[12] == pytest.approx(U * V)
    assert abs(t[33]) < 1e-15 and abs(t[13]) < 1e-15 and abs(t[23]) < 1e-15


# ----------------------------------------------------------------------------- KE / TKE known answers
@pytest.mark.parametrize("gname", GRIDS)
@pytest.mark.parametrize("cname", ["scalar", "smagorinsky_like", "tuple"])
def test_ke_dissipation_of_linear_shear(gname, cname):
    """test/test_turbulent_kinetic_energy_equation.jl:120-145: ε = ν (du/dz)², rtol 1e-12 on the
    regular grid / 0.06 on the stretched one; zero about the horizontal mean."""
    g = GRIDS[gname]()
    cl = make_closures(g, np.random.default_rng(1))[cname]
    dudz = 2.0
    u, v, w = velocities(g, lambda x, y, z: dudz * z)
    eps = K.evaluate(K.viscous_dissipation_rate_ccc, g, "ccc", None, dict(u=u, v=v, w=w), dict(closure=cl))
    rtol = 1e-12 if gname == "regular" else 0.06
    if cname == "scalar":
        assert eps[IDX] == pytest.approx(nu_ccc(g, cl, IDX) * dudz ** 2, rel=rtol)
    Ubar = average_dims(u, (1, 2))
    up = materialize(g, "fcc", u.interior - Ubar.interior)
    epsp = K.evaluate(K.viscous_dissipation_rate_ccc, g, "ccc", None, dict(u=up, v=v, w=w), dict(closure=cl))
    assert abs(epsp[IDX]) <= np.finfo(float).eps


def test_ke_advection_known_answer():
    """test/test_kinetic_energy_equation.jl:41-55: u = C1 y, v = C2  =>  ADV = C1² C2 y."""
    g = regular_grid()
    C1, C2 = 2.0, 3.0
    u, v, w = velocities(g, lambda x, y, z: C1 * y, C2)
    adv = K.evaluate(K.ke_advection_ccc, g, "ccc", dict(u=u, v=v, w=w))
    yc = g.centers[1][np.arange(2, N)]
    np.testing.assert_allclose(adv[0, 1:N - 1, 0], C1 ** 2 * C2 * yc, rtol=1e-12)


def test_centered_fourth_order_advection():
    """Centered(order = 4) -- the scheme of every example of the reference (docs/examples/*.jl: `advection = Centered(order=4)`).
    (a) the reference's own advection known answer u = C1 y, v = C2 => ADV = C1² C2 y (test/test_kinetic_energy_equation.jl:41-55)
    holds for any consistent centred scheme on these linear fields; (b) with a constant advecting velocity the flux divergence of
    the (-1, 7, 7, -1)/12 interpolation IS the classic fourth-order central difference: the error of div_Uc on c = sin 2πx drops 16 x
    per halving of the spacing (4 x with order 2) -- this pins the coefficients and the flux positions; (c) along the Bounded z axis
    the stencil falls back to order 2 next to the walls [OCN-mem: outside_symmetric_halo], and only there."""
    g = regular_grid()
    C1, C2 = 2.0, 3.0
    u, v, w = velocities(g, lambda x, y, z: C1 * y, C2)
    adv = K.evaluate(K.ke_advection_ccc, g, "ccc", dict(u=u, v=v, w=w), "Centered(order=4)")
    yc = g.centers[1][np.arange(3, N - 1)]
    np.testing.assert_allclose(adv[0, 2:N - 2, 0], C1 ** 2 * C2 * yc, rtol=1e-12)      # rows whose stencil stays clear of the periodic seam
    errs = {}
    for order in ("Centered", "Centered(order=4)"):
        for n in (16, 32):
            gg = OGrid((n, 4, 4), extent=(1, 1, 1))
            uu = OField(gg, "fcc").set(1.5).fill_halo()
            vv = OField(gg, "cfc").set(0.0).fill_halo()
            ww = OField(gg, "ccf").set(0.0).fill_halo()
            c = OField(gg, "ccc").set(lambda x, y, z: np.sin(2 * np.pi * x)).fill_halo()
            got = K.evaluate(K.div_Uc, gg, "ccc", order, (uu, vv, ww), c)[:, 0, 1]
            xc = (np.arange(n) + 0.5) / n
            errs[(order, n)] = np.abs(got - 1.5 * 2 * np.pi * np.cos(2 * np.pi * xc)).max()
    assert 3.8 < errs[("Centered", 16)] / errs[("Centered", 32)] < 4.2
    assert 15.0 < errs[("Centered(order=4)", 16)] / errs[("Centered(order=4)", 32)] < 17.0
    assert errs[("Centered(order=4)", 32)] < 1e-2 * errs[("Centered", 32)]
    # (c) vertical advection of a tracer by w = const: faces 2 and Nz (and the wall faces) are second order, faces 3 .. Nz-1 fourth
    gz = OGrid((4, 4, 8), extent=(1, 1, 1))
    zero_u, zero_v = OField(gz, "fcc").set(0.0).fill_halo(), OField(gz, "cfc").set(0.0).fill_halo()
    wz = OField(gz, "ccf").set(1.0).fill_halo(impenetrable=False)
    cz = OField(gz, "ccc").set(lambda x, y, z: z ** 3).fill_halo()
    kk = np.arange(1, 10)[None, None, :]
    one = np.array([[[1]]])
    f2 = K._adv_flux_z(one, one, kk, gz, 2, (zero_u, zero_v, wz), cz)[0, 0]
    f4 = K._adv_flux_z(one, one, kk, gz, 4, (zero_u, zero_v, wz), cz)[0, 0]
    same = np.isclose(f2, f4, rtol=1e-14, atol=0)
    assert same.tolist() == [True, True, False, False, False, False, False, True, True]


@pytest.mark.parametrize("gname", GRIDS)
def test_buoyancy_production_known_answer(gname):
    """test/test_kinetic_energy_equation.jl:131-140: w = w0, b = b0  =>  wb == w0 b0 exactly."""
    g = GRIDS[gname]()
    u, v, _ = velocities(g)
    w = OField(g, "ccf").set(2.0).fill_halo(impenetrable=False)   # enforce_incompressibility=false
    b = OField(g, "ccc").set(3.0).fill_halo()
    wb = K.evaluate(K.ke_buoyancy_ccc, g, "ccc", dict(u=u, v=v, w=w), (0, 0, 1), b)
    assert wb[0, 0, 1] == 6.0


@pytest.mark.parametrize("gname", GRIDS)
def test_total_shear_production_is_sum(gname):
    """test/test_turbulent_kinetic_energy_equation.jl:87-107."""
    g = GRIDS[gname]()
    u, v, w = velocities(g, lambda x, y, z: np.sin(2 * np.pi * x) + 0.1 * np.cos(2 * np.pi * y),
                         lambda x, y, z: np.cos(2 * np.pi * x) + 0.1 * np.sin(2 * np.pi * z))
    U, V, W = (average_dims(f, (2, 3)) for f in (u, v, w))
    args = (DiffField(u, U), DiffField(v, V), DiffField(w, W), U, V, W)
    x = K.evaluate(K.shear_production_rate_x_ccc, g, "ccc", *args)
    y = K.evaluate(K.shear_production_rate_y_ccc, g, "ccc", *args)
    z = K.evaluate(K.shear_production_rate_z_ccc, g, "ccc", *args)
    tot = K.evaluate(K.shear_production_rate_ccc, g, "ccc", *args)
    assert np.any(tot != 0)
    np.testing.assert_allclose(tot, x + y + z, rtol=1e-12, atol=1e-14)
    tke = K.evaluate(K.turbulent_kinetic_energy_ccc, g, "ccc", u, v, w, U, V, W)
    assert tke.min() >= 0


@pytest.mark.parametrize("gname", GRIDS)
@pytest.mark.parametrize("cname", ["scalar", "smagorinsky_like", "tuple"])
def test_tracer_variance_integral_identity(gname, cname):
    """test/test_tracer_variance_diagnostics.jl:31-34: <χ> == <2c∇·F>, rtol 1e-12."""
    g = GRIDS[gname]()
    rng = np.random.default_rng(2)
    cl = make_closures(g, rng)[cname]
    b = OField(g, "ccc")
    b.interior[...] = rng.standard_normal(b.interior.shape)
    b.fill_halo()
    chi = K.evaluate(K.tracer_variance_dissipation_rate_ccc, g, "ccc", cl, None, 1, b)
    dif = K.evaluate(K.c_div_q_c, g, "ccc", cl, None, 1, b)
    a1, a2 = K.average(chi, g, "ccc"), K.average(dif, g, "ccc")
    assert a1 == pytest.approx(a2, rel=1e-12, abs=2 * np.finfo(float).eps)


@pytest.mark.parametrize("gname", GRIDS)
@pytest.mark.parametrize("cname", ["scalar", "smagorinsky_like", "tuple"])
def test_ke_stress_integral_equals_the_dissipation_integral(gname, cname):
    """Summation by parts of DIFF = uᵢ∂ⱼτᵢⱼ (src/KineticEnergyEquation.jl:191-202) over a periodic x-y domain with impenetrable,
    no-flux top and bottom (w = 0 on the boundary faces, mirrored u and v: every boundary flux τ₁₃, τ₂₃ vanishes) gives
    <uᵢ∂ⱼτᵢⱼ> = -<τᵢⱼ∂ⱼuᵢ> = <ε> (src/KineticEnergyEquation.jl:441-453): the closing relation of the reference's KE budget
    (docs/examples/two_dimensional_turbulence.jl), here between two independently restated kernels."""
    g = GRIDS[gname]()
    rng = np.random.default_rng(5)
    cl = make_closures(g, rng)[cname]
    f = {}
    for name, loc in (("u", "fcc"), ("v", "cfc"), ("w", "ccf")):
        fld = OField(g, loc)
        fld.interior[...] = rng.standard_normal(fld.interior.shape)
        f[name] = fld.fill_halo()
    eps = K.evaluate(K.viscous_dissipation_rate_ccc, g, "ccc", None, f, {"closure": cl})
    dif = K.evaluate(K.ke_stress_ccc, g, "ccc", cl, None, None, f, None)
    assert K.integral(dif, g, "ccc") == pytest.approx(K.integral(eps, g, "ccc"), rel=1e-11)
    assert not np.allclose(dif, eps)            # equal in the integral only


@pytest.mark.parametrize("gname", GRIDS)
def test_ke_forcing_known_answer(gname):
    """test/test_kinetic_energy_equation.jl:105-129: Fᵘ = -u, Fᵛ = -v, Fʷ = -w on noise  =>
    uᵢFᵤᵢ = @at ccc (-u² - v² - w²), rtol 1e-12."""
    g = GRIDS[gname]()
    rng = np.random.default_rng(9)
    f = {}
    for name, loc in (("u", "fcc"), ("v", "cfc"), ("w", "ccf")):
        fld = OField(g, loc)
        fld.interior[...] = rng.standard_normal(fld.interior.shape)
        f[name] = fld.fill_halo()
    damp = lambda x, y, z, t, q: -q
    forcings = {"u": K.ContinuousForcing(damp, "fcc", ("u",)), "v": K.ContinuousForcing(damp, "cfc", ("v",)),
                "w": K.ContinuousForcing(damp, "ccf", ("w",))}
    got = K.evaluate(K.ke_forcing_ccc, g, "ccc", forcings, None, f)
    sq = lambda i, j, k, gg, q: -q[i, j, k] ** 2
    truth = (K.evaluate(K.ix_caa, g, "ccc", sq, f["u"]) + K.evaluate(K.iy_aca, g, "ccc", sq, f["v"]) +
             K.evaluate(K.iz_aac, g, "ccc", sq, f["w"]))
    np.testing.assert_allclose(got, truth, rtol=1e-12, atol=np.finfo(float).eps)
    # no forcing at all: zero
    assert not K.evaluate(K.ke_forcing_ccc, g, "ccc", {}, None, f).any()


def _tendency_args(cl, vel, b, forcing=None):
    # (advection, closure, c_immersed_bc, buoyancy, biogeochemistry, background_fields, velocities, tracers, auxiliary_fields,
    #  closure_fields, clock, forcing): the dependency tuple of src/TracerVarianceEquation.jl:72-85
    return (None, cl, None, None, None, None, vel, {"b": b}, None, None, None, forcing)


@pytest.mark.parametrize("gname", GRIDS)
@pytest.mark.parametrize("cname", ["scalar", "smagorinsky_like", "tuple"])
def test_tracer_variance_tendency_limits(gname, cname):
    """c∂ₜcᶜᶜᶜ = 2c (-div_Uc - ∇·q + F) (src/TracerVarianceEquation.jl:18-39) in three limits: at rest and unforced it is
    -c∇_dot_qᶜ (:93-98) cell by cell; a uniform horizontal flow moves variance around without creating any, so
    <2c∂ₜc> = -<χ> (:157-161); and a pure relaxation F = -r c without flow or diffusion gives -2 r c²."""
    g = GRIDS[gname]()
    rng = np.random.default_rng(11)
    cl = make_closures(g, rng)[cname]
    b = OField(g, "ccc")
    b.interior[...] = rng.standard_normal(b.interior.shape)
    b.fill_halo()

    def vel(u0, v0):
        out = {}
        for name, loc, val in (("u", "fcc", u0), ("v", "cfc", v0), ("w", "ccf", 0.0)):
            f = OField(g, loc)
            f.interior[...] = val
            out[name] = f.fill_halo()
        return out
    rest = K.evaluate(K.c_dt_c_ccc, g, "ccc", 1, "b", *_tendency_args(cl, vel(0.0, 0.0), b))
    dif = K.evaluate(K.c_div_q_c, g, "ccc", cl, None, 1, b)
    np.testing.assert_allclose(rest, -dif, rtol=1e-13, atol=1e-300)
    moving = K.evaluate(K.c_dt_c_ccc, g, "ccc", 1, "b", *_tendency_args(cl, vel(0.7, -0.4), b))
    chi = K.evaluate(K.tracer_variance_dissipation_rate_ccc, g, "ccc", cl, None, 1, b)
    assert not np.allclose(moving, rest)
    assert K.integral(moving, g, "ccc") == pytest.approx(-K.integral(chi, g, "ccc"), rel=1e-10)
    relax = K.ContinuousForcing(lambda x, y, z, t, c: -0.3 * c, "ccc", ("b",))
    forced = K.evaluate(K.c_dt_c_ccc, g, "ccc", 1, "b", *_tendency_args(K.Closure(nu=0.0, kappa=0.0), vel(0.0, 0.0), b, relax))
    np.testing.assert_allclose(forced, -0.6 * b.interior ** 2, rtol=1e-13, atol=1e-300)


def _random_state(g, seed):
    rng = np.random.default_rng(seed)
    f = {}
    for name, loc in (("u", "fcc"), ("v", "cfc"), ("w", "ccf"), ("b", "ccc"), ("p", "ccc")):
        fld = OField(g, loc)
        fld.interior[...] = rng.standard_normal(fld.interior.shape)
        f[name] = fld.fill_halo()
    return f


def _ke_tendency(g, cl, f, coriolis=None, buoyancy=None, pHY=None, forcings=None):
    vel = {"u": f["u"], "v": f["v"], "w": f["w"]}
    return K.evaluate(K.ke_tendency_ccc, g, "ccc", None, coriolis, None, cl, None, None, None, buoyancy, None, vel, {"b": f["b"]}, None,
                      None, pHY, None, forcings)


@pytest.mark.parametrize("gname", GRIDS)
@pytest.mark.parametrize("cname", ["scalar", "smagorinsky_like", "tuple"])
def test_ke_tendency_is_the_sum_of_the_pinned_budget_terms(gname, cname):
    """uᵢGᵢ (src/KineticEnergyEquation.jl:74-95) is linear in the velocity tendency, and every piece of Oceananigans'
    NonhydrostaticModel tendency has its own Oceanostics diagnostic that is pinned above (ADV = C₁²C₂y, wb = w₀b₀, <DIFF> = <ε>,
    uᵢFᵢ = -Σuᵢ²): cell by cell  uᵢGᵢ = -ADV + uᵢbᵢ - uᵢ(f x u)ᵢ - uₕ·∇ₕpHY′ - uᵢ∂ⱼτᵢⱼ + uᵢFᵢ  (KET + PR closes the budget:
    :100-103).  With a hydrostatic pressure anomaly the vertical buoyancy term leaves the tendency."""
    g = GRIDS[gname]()
    cl = make_closures(g, np.random.default_rng(3))[cname]
    f = _random_state(g, 21)
    vel = {"u": f["u"], "v": f["v"], "w": f["w"]}
    damp = lambda x, y, z, t, q: -0.3 * q
    forcings = {"u": K.ContinuousForcing(damp, "fcc", ("u",)), "v": K.ContinuousForcing(damp, "cfc", ("v",)),
                "w": K.ContinuousForcing(damp, "ccf", ("w",))}
    fcor, ghat = (1e-2, -2e-2, 0.7), (0.0, 0.0, 1.0)
    ev = K.evaluate
    adv = ev(K.ke_advection_ccc, g, "ccc", vel)
    wb = ev(K.ke_buoyancy_ccc, g, "ccc", vel, ghat, f["b"])
    dif = ev(K.ke_stress_ccc, g, "ccc", cl, None, None, {**vel, "b": f["b"]}, None)
    frc = ev(K.ke_forcing_ccc, g, "ccc", forcings, None, {**vel, "b": f["b"]})
    cor = (ev(K.ix_caa, g, "ccc", K.psif, f["u"], K.x_f_cross_U, fcor, vel) + ev(K.iy_aca, g, "ccc", K.psif, f["v"], K.y_f_cross_U, fcor, vel) +
           ev(K.iz_aac, g, "ccc", K.psif, f["w"], K.z_f_cross_U, fcor, vel))
    got = _ke_tendency(g, cl, f, coriolis=fcor, buoyancy=(ghat, "b"), forcings=forcings)
    want = -adv + wb - cor - dif + frc
    scale = np.abs(adv).max() + np.abs(dif).max() + np.abs(wb).max()
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-12 * scale)
    assert np.abs(cor).max() > 1e-3 and np.abs(got).max() > 1e-2
    # hydrostatic split: w loses its buoyancy term, u and v feel -∇ₕ pHY′
    zero = OField(g, "ccf").set(0.0).fill_halo()
    wb_h = ev(K.ke_buoyancy_ccc, g, "ccc", {"u": f["u"], "v": f["v"], "w": zero}, ghat, f["b"])     # horizontal part (zero here)
    pr_h = ev(K.ke_pressure_ccc, g, "ccc", {"u": f["u"], "v": f["v"], "w": zero}, f["p"])
    got_h = _ke_tendency(g, cl, f, coriolis=fcor, buoyancy=(ghat, "b"), pHY=f["p"], forcings=forcings)
    np.testing.assert_allclose(got_h, -adv + wb_h - cor - pr_h - dif + frc, rtol=0, atol=1e-12 * (scale + np.abs(pr_h).max()))


def test_ke_tendency_of_uniform_flow_on_an_f_plane_vanishes():
    """A uniform horizontal flow has no advection, stress or pressure tendency, and the Coriolis force does no work:
    x_f_cross_U = -f v₀, y_f_cross_U = f u₀  =>  uᵢGᵢ = u₀ f v₀ - v₀ f u₀ = 0, while each part is f u₀ v₀."""
    g = regular_grid()
    u, v, w = velocities(g, 2.0, 3.0, 0.0)
    b = OField(g, "ccc").set(0.0).fill_halo()
    f = {"u": u, "v": v, "w": w, "b": b}
    got = _ke_tendency(g, K.Closure(nu=1e-3, kappa=1e-3), f, coriolis=(0.0, 0.0, 0.5))
    assert np.abs(got).max() <= 1e-15
    vel = {"u": u, "v": v, "w": w}
    part = K.evaluate(K.ix_caa, g, "ccc", K.psif, u, K.x_f_cross_U, (0.0, 0.0, 0.5), vel)
    np.testing.assert_allclose(part, -0.5 * 2.0 * 3.0, rtol=1e-14)


@pytest.mark.parametrize("gname", GRIDS)
def test_smagorinsky_viscosity_of_a_linear_shear(gname):
    """u = S z: ΣᵢⱼΣᵢⱼ = S²/2 (test/test_canonical_flows.jl:130-146 pins |S|), so νₑ = (C Δᶠ)² S with Δᶠ = ∛(ΔxΔyΔz); on a
    linear stratification b = N² z the Lilly factor is ς = √(1 - Cb N²/(Pr S²/2)), zero once Cb N² ≥ Pr S²/2."""
    g = GRIDS[gname]()
    S, N2, C = 2.0, 0.5, 0.16
    u, v, w = velocities(g, lambda x, y, z: S * z)
    b = OField(g, "ccc").set(lambda x, y, z: N2 * z).fill_halo()
    i, j, k = IDX
    delta = np.cbrt(g.Dx("ccc", i + 1, j + 1, k + 1) * g.Dy("ccc", i + 1, j + 1, k + 1) * g.Dz("ccc", i + 1, j + 1, k + 1))
    rtol = 1e-12 if gname == "regular" else 0.06          # the stretched grid's tolerance for shear diagnostics (as for ε above)
    nu = K.evaluate(K.smagorinsky_nu_ccc, g, "ccc", u, v, w, b, C)
    assert nu[IDX] == pytest.approx((C * delta) ** 2 * S, rel=rtol)
    nu_b = K.evaluate(K.smagorinsky_nu_ccc, g, "ccc", u, v, w, b, C, 1.0, 1.0)
    if gname == "regular":
        assert nu_b[IDX] == pytest.approx((C * delta) ** 2 * S * np.sqrt(1 - N2 / (S ** 2 / 2)), rel=1e-12)
    assert np.all(nu_b <= nu + 1e-18)
    nu_stable = K.evaluate(K.smagorinsky_nu_ccc, g, "ccc", u, v, w, b, C, 100.0, 1.0)
    assert nu_stable[IDX] == 0.0
    # unstable stratification is clipped (N²⁺ = max(0, N²)): no enhancement
    b_un = OField(g, "ccc").set(lambda x, y, z: -N2 * z).fill_halo()
    np.testing.assert_array_equal(K.evaluate(K.smagorinsky_nu_ccc, g, "ccc", u, v, w, b_un, C, 1.0, 1.0), nu)


def test_richardson_and_ertel_pv_known_answers():
    """test/test_active_tracer_diagnostics.jl:24-80: u = S z + S y, b = N² z  =>
    Ri[3,3,3] = N²/S², EPV[3,3,3] = N² (f_z - S)."""
    g = regular_grid()
    N2, S, fz = 1e-5, 1e-3, 1e-4
    u, v, w = velocities(g, lambda x, y, z: S * z + S * y)
    b = OField(g, "ccc").set(lambda x, y, z: N2 * z).fill_halo()
    Ri = K.evaluate(K.richardson_number_ccf, g, "ccf", u, v, w, b, (0, 0, 1))
    assert Ri[2, 2, 2] == pytest.approx(N2 / S ** 2)
    epv = K.evaluate(K.ertel_potential_vorticity_fff, g, "fff", u, v, w, b, 0, 0, fz)
    assert epv[2, 2, 2] == pytest.approx(N2 * (fz - S))
    depv = K.evaluate(K.directional_ertel_potential_vorticity_fff, g, "fff", u, v, w, b,
                      dict(f_dir=fz, dir_x=0, dir_y=0, dir_z=1))
    assert depv[2, 2, 2] == pytest.approx(N2 * (fz - S))


def test_ertel_pv_doctest_golden_column():
    """src/FlowDiagnostics.jl:288-327 (doctest): Flat/Flat/Bounded 4-cell column, GradientBC(N²) on top."""
    g = OGrid(4, extent=1, topology=("F", "F", "B"))
    N2 = 1e-6
    b = OField(g, "ccc").set(lambda x, y, z: N2 * z).fill_halo(top_gradient=N2)
    u, v, w = velocities(g)
    epv = K.evaluate(K.ertel_potential_vorticity_fff, g, "fff", u, v, w, b, 0, 0, 1e-4).ravel()
    golden = np.array([0.0, 1.0000000000000002e-10, 9.999999999999998e-11, 1.0000000000000002e-10, 1.0e-10])
    np.testing.assert_allclose(epv, golden, rtol=1e-12, atol=0)
    assert epv[0] == 0.0


# ----------------------------------------------------------------------------- filters
def make_grid(Nx=8, Ny=8, Nz=8, halo=(2, 2, 2), topology=("P", "P", "P")):      # test/test_spatial_filters.jl:18-23
    return OGrid((Nx, Ny, Nz), x=(0, 1), y=(0, 1), z=(0, 1), halo=halo, topology=topology)


def ref_weighted(ic, dims, width, Ns, w1d, mode, left=None, right=None):
    """Independent brute-force N-D references, test/test_spatial_filters.jl:47-142."""
    Nx, Ny, Nz = Ns
    rng = lambda d: range(-width, width + 1) if d in dims else range(0, 1)
    ref = np.empty_like(ic)
    for i in range(Nx):
        for j in range(Ny):
            for k in range(Nz):
                s = 0.0
                ws = 0.0
                for di in rng(1):
                    for dj in rng(2):
                        for dk in rng(3):
                            idx = [i + di, j + dj, k + dk]
                            w = 1.0
                            for d, dd in zip((1, 2, 3), (di, dj, dk)):
                                if d in dims:
                                    w *= w1d[dd + width]
                            ok = True
                            val = None
                            for a, d in enumerate((1, 2, 3)):
                                if d not in dims:
                                    continue
                                n = Ns[a]
                                if mode == "periodic":
                                    idx[a] %= n
                                elif mode == "edge":
                                    idx[a] = min(max(idx[a], 0), n - 1)
                                elif mode == "shrink":
                                    ok = ok and 0 <= idx[a] < n
                                elif mode == "constant":
                                    if idx[a] < 0:
                                        val = left
                                    elif idx[a] >= n:
                                        val = right
                            if not ok:
                                continue
                            s += w * (ic[tuple(idx)] if val is None else val)
                            ws += w
                ref[i, j, k] = s / ws
    return ref


box_w = lambda w: [1.0] * (2 * w + 1)
gauss_w = lambda w, s=1.0: [np.exp(-(i - w) ** 2 / (2 * s ** 2)) for i in range(2 * w + 1)]
FILTERS = {"box": (lambda f, dims, width, **kw: OF.box_filter(f, dims, 2 * width + 1, **kw), box_w),
           "gauss": (lambda f, dims, width, **kw: OF.gaussian_filter(f, dims, 1 / 8, 2 * width + 1, **kw), gauss_w)}


def center_field_from(g, f, loc="ccc"):
    return OField(g, loc).set(f).fill_halo()


@pytest.mark.parametrize("fname", FILTERS)
def test_filter_periodic_stencil_sum(fname):
    """test/test_spatial_filters.jl:169-181."""
    filt, mkw = FILTERS[fname]
    g = make_grid()
    c = center_field_from(g, lambda x, y, z: np.sin(2 * np.pi * x) * np.cos(2 * np.pi * y) + z ** 2)
    ic = c.interior.copy()
    for dims, width in [((1,), 1), ((1, 2, 3), 1), ((1, 2), 2), ((1, 2, 3), 2)]:
        np.testing.assert_allclose(filt(c, dims, width), ref_weighted(ic, dims, width, (8, 8, 8), mkw(width), "periodic"), rtol=1e-12, atol=1e-14)


@pytest.mark.parametrize("fname", FILTERS)
def test_filter_bounded_policies(fname):
    """test/test_spatial_filters.jl:280-330 (shrink / edge / constant pad / per-dim)."""
    filt, mkw = FILTERS[fname]
    g = make_grid(halo=(1, 1, 1), topology=("B", "B", "B"))
    c = center_field_from(g, lambda x, y, z: np.sin(2 * np.pi * x) * np.cos(2 * np.pi * y) + z ** 2)
    ic = c.interior.copy()
    Ns = (8, 8, 8)
    for dims in [(1,), (2,), (3,), (1, 2), (1, 2, 3)]:
        np.testing.assert_allclose(filt(c, dims, 2), ref_weighted(ic, dims, 2, Ns, mkw(2), "shrink"), rtol=1e-12, atol=1e-14)
    for dims in [(1,), (2,), (3,), (1, 2, 3)]:
        np.testing.assert_allclose(filt(c, dims, 2, boundary="edge"), ref_weighted(ic, dims, 2, Ns, mkw(2), "edge"), rtol=1e-12, atol=1e-14)
    c2 = center_field_from(g, lambda x, y, z: np.sin(2 * np.pi * x) * np.cos(2 * np.pi * y))
    for d, width in [(1, 2), (2, 3), (3, 2)]:
        out = filt(c2, (d,), width, boundary=dict(left=-1.25, right=2.5))
        np.testing.assert_allclose(out, ref_weighted(c2.interior.copy(), (d,), width, Ns, mkw(width), "constant", -1.25, 2.5), rtol=1e-12, atol=1e-14)
    out = filt(c2, (1, 2), 2, boundary=("shrink", "edge"))
    sx = ref_weighted(c2.interior.copy(), (1,), 2, Ns, mkw(2), "shrink")
    np.testing.assert_allclose(out, ref_weighted(sx, (2,), 2, Ns, mkw(2), "edge"), rtol=1e-12, atol=1e-14)


@pytest.mark.parametrize("fname", FILTERS)
def test_filter_face_location_bounded(fname):
    """test/test_spatial_filters.jl:337-359: a ZFaceField on Bounded z has Nz+1 faces in range."""
    filt, mkw = FILTERS[fname]
    g = make_grid(topology=("P", "P", "B"))
    w = OField(g, "ccf").set(lambda x, y, z: np.sin(2 * np.pi * x) + z ** 2).fill_halo(impenetrable=False)
    iw = w.interior.copy()
    assert iw.shape == (8, 8, 9)
    for boundary in ("shrink", "edge"):
        np.testing.assert_allclose(filt(w, (3,), 2, boundary=boundary), ref_weighted(iw, (3,), 2, (8, 8, 9), mkw(2), boundary), rtol=1e-12, atol=1e-14)


def test_box_filter_hand_computed_vectors():
    """test/test_spatial_filters.jl:381-431 (golden vectors)."""
    g = OGrid(10, x=(0, 1), halo=2, topology=("P", "F", "F"))
    c = OField(g, "ccc")
    c.interior[:, 0, 0] = np.arange(10.0)
    np.testing.assert_allclose(OF.box_filter(c, (1,), 3).ravel(), [10 / 3, 1, 2, 3, 4, 5, 6, 7, 8, 17 / 3], rtol=1e-14)
    np.testing.assert_allclose(OF.box_filter(c, (1,), 5).ravel(), [4, 3, 2, 3, 4, 5, 6, 7, 6, 5], rtol=1e-14)
    g = OGrid(10, x=(0, 1), halo=1, topology=("B", "F", "F"))
    c = OField(g, "ccc")
    c.interior[:, 0, 0] = np.arange(10.0)
    chk = lambda N, exp, **kw: np.testing.assert_allclose(OF.box_filter(c, (1,), N, **kw).ravel(), exp, rtol=1e-14)
    chk(5, [1, 1.5, 2, 3, 4, 5, 6, 7, 7.5, 8])
    chk(7, [1.5, 2, 2.5, 3, 4, 5, 6, 6.5, 7, 7.5])
    chk(5, [0.6, 1.2, 2, 3, 4, 5, 6, 7, 7.8, 8.4], boundary="edge")
    chk(7, [6 / 7, 10 / 7, 15 / 7, 3, 4, 5, 6, 48 / 7, 53 / 7, 57 / 7], boundary="edge")
    chk(5, [0.2, 1, 2, 3, 4, 5, 6, 7, 5.6, 4], boundary=dict(left=-1.0, right=-2.0))
    chk(7, [3 / 7, 8 / 7, 2, 3, 4, 5, 6, 37 / 7, 31 / 7, 24 / 7], boundary=dict(left=-1.0, right=-2.0))


def test_gaussian_dirac_delta_impulse_response():
    """test/test_spatial_filters.jl:714-745."""
    Nn = 21
    g = OGrid(Nn, x=(0, 1), halo=3, topology=("P", "F", "F"))
    i0 = Nn // 2
    for width, sc in [(3, 1.0), (3, 2.0), (5, 1.5)]:
        c = OField(g, "ccc")
        c.interior[i0, 0, 0] = 1
        out = OF.gaussian_filter(c, (1,), sc / Nn, 2 * width + 1).ravel()
        wts = gauss_w(width, sc)
        exp = np.zeros(Nn)
        for j in range(Nn):
            if abs(i0 - j) <= width:
                exp[j] = wts[i0 - j + width] / sum(wts)
        np.testing.assert_allclose(out, exp, rtol=1e-12, atol=1e-300)


def test_gaussian_identity_on_face_field_bit_exact():
    """test/test_spatial_filters.jl:701-712: σ << Δ, N=3  =>  output == input bit for bit, incl. last face."""
    g = make_grid(topology=("P", "P", "B"))
    w = OField(g, "ccf").set(lambda x, y, z: np.sin(2 * np.pi * x) + z ** 2).fill_halo(impenetrable=False)
    for dims in ((3,), (1, 3), (1, 2, 3)):
        out = OF.gaussian_filter(w, dims, 1e-4, 3)
        assert np.all(np.isfinite(out)) and np.array_equal(out, w.interior)


def test_gaussian_width_inference_and_tuple_order():
    """test/test_spatial_filters.jl:445-506."""
    g = make_grid()
    c = center_field_from(g, lambda x, y, z: np.sin(2 * np.pi * x) + np.cos(2 * np.pi * z))
    s = 1.5 / 8
    assert np.array_equal(OF.gaussian_filter(c, (1,), s), OF.gaussian_filter(c, (1,), s, 7))
    assert np.array_equal(OF.gaussian_filter(c, (1, 3), s), OF.gaussian_filter(c, (1, 3), s, 7))
    assert np.array_equal(OF.gaussian_filter(c, (3, 1), s, (5, 7)), OF.gaussian_filter(c, (1, 3), s, (7, 5)))
    ga = OGrid((8, 8, 8), x=(0, 1), y=(0, 1), z=(0, 2), topology=("P", "P", "P"))
    ca = center_field_from(ga, lambda x, y, z: np.sin(2 * np.pi * x) * np.cos(np.pi * z))
    assert np.array_equal(OF.gaussian_filter(ca, (1, 3), 0.2), OF.gaussian_filter(ca, (1, 3), 0.2, (9, 5)))
    for bad in [(5,), (5, 7, 9), (3, 4), (5, 1), (5, 3.5)]:
        with pytest.raises(ValueError):
            OF.gaussian_filter(c, (1, 3), s, bad)


def ref_gaussian_stretched(c, d, sigma, width, boundary="shrink"):
    """test/test_spatial_filters.jl:557-591."""
    g = c.grid
    n = c.ext[d]
    idx = np.arange(1, n + 1)
    coords, spac = g.node(d, c.loc[d], idx), g.spacing(d, c.loc[d], idx)
    L = g.L[d]
    periodic = g.topology[d] == "P"
    ic = c.interior
    ref = np.empty_like(ic)
    for I in np.ndindex(ic.shape):
        ctr = I[d]
        x0 = coords[ctr]
        s = ws = 0.0
        for D in range(-width, width + 1):
            m = ctr + D
            J = list(I)
            if periodic:
                mr = m % n
                pos = coords[mr] - L * (m < 0) + L * (m >= n)
                J[d] = mr
                val, cnt = ic[tuple(J)], 1
            else:
                mr = min(max(m, 0), n - 1)
                pos = coords[mr]
                J[d] = mr
                inb = 0 <= m < n
                if boundary == "shrink":
                    val, cnt = (ic[tuple(J)] if inb else 0.0), int(inb)
                else:
                    val, cnt = ic[tuple(J)], 1
            w = spac[mr] * np.exp(-(pos - x0) ** 2 / (2 * sigma ** 2))
            s += w * val
            ws += w * cnt
        ref[I] = s / ws
    return ref


def test_gaussian_stretched_numerical():
    """test/test_spatial_filters.jl:595-616."""
    zf = lambda k: -1 + (k - 1) / 8 + 0.08 * np.sin(2 * np.pi * (k - 1) / 8)
    gz = OGrid((6, 6, 8), x=(0, 1), y=(0, 1), z=zf, halo=(3, 3, 3), topology=("P", "P", "B"))
    cz = center_field_from(gz, lambda x, y, z: np.sin(2 * np.pi * x) * np.cos(2 * np.pi * y) + z ** 2 + 0.3 * z)
    xf = lambda i: (i - 1) / 8 + 0.1 * np.sin(2 * np.pi * (i - 1) / 8)
    gx = OGrid((8, 4, 4), x=xf, y=(0, 1), z=(0, 1), halo=(3, 1, 1), topology=("P", "P", "B"))
    cx = center_field_from(gx, lambda x, y, z: np.sin(2 * np.pi * x))
    for width in (2, 3):
        for boundary in ("shrink", "edge"):
            out = OF.gaussian_filter(cz, (3,), 0.15, 2 * width + 1, boundary=boundary)
            np.testing.assert_allclose(out, ref_gaussian_stretched(cz, 2, 0.15, width, boundary), rtol=1e-12, atol=1e-14)
        out = OF.gaussian_filter(cx, (1,), 0.12, 2 * width + 1)
        np.testing.assert_allclose(out, ref_gaussian_stretched(cx, 0, 0.12, width), rtol=1e-12, atol=1e-14)


def test_gaussian_stretched_dirac_delta():
    """test/test_spatial_filters.jl:658-689."""
    Nn = 21
    zf = lambda k: (k - 1) / Nn + 0.06 * np.sin(2 * np.pi * (k - 1) / Nn)
    g = OGrid((1, 1, Nn), x=(0, 1), y=(0, 1), z=zf, halo=(1, 1, 4), topology=("P", "P", "B"))
    idx = np.arange(1, Nn + 1)
    zc, dz = g.node(2, "c", idx), g.spacing(2, "c", idx)
    i0 = Nn // 2
    for width, sig in [(3, 0.05), (4, 0.07)]:
        c = OField(g, "ccc")
        c.interior[0, 0, i0] = 1
        out = OF.gaussian_filter(c, (3,), sig, 2 * width + 1).ravel()
        exp = np.zeros(Nn)
        for k in range(Nn):
            ws = w0 = 0.0
            for D in range(-width, width + 1):
                m = k + D
                if 0 <= m < Nn:
                    w = dz[m] * np.exp(-(zc[m] - zc[k]) ** 2 / (2 * sig ** 2))
                    ws += w
                    if m == i0:
                        w0 = w
            exp[k] = w0 / ws
        np.testing.assert_allclose(out, exp, rtol=1e-12, atol=1e-300)


def test_filter_validation_errors():
    """test/test_spatial_filters.jl:224-265, 369-377."""
    g = make_grid()
    c = OField(g, "ccc")
    for dims in [(), (0,), (4,), (1, 1), ("x",), [1, 2]]:
        with pytest.raises(ValueError):
            OF.box_filter(c, dims, 3)
    for Nbad in [0, -1, 1, 2, 4, 3.5]:
        with pytest.raises(ValueError):
            OF.box_filter(c, (1,), Nbad)
    small = OField(make_grid(4, 4, 4), "ccc")
    with pytest.raises(ValueError):
        OF.box_filter(small, (1,), 11)
    OF.box_filter(small, (1,), 9)
    OF.box_filter(OField(make_grid(4, 4, 4, topology=("B", "B", "B")), "ccc"), (1,), 11)
    for b in ["foo", dict(foo=1), dict(left=1, right=2, mid=3), 42, ("shrink",)]:
        with pytest.raises(ValueError):
            OF.box_filter(c, (1, 2) if isinstance(b, tuple) else (1,), 3, boundary=b)
    for s in [0, -1.0, "foo"]:
        with pytest.raises(ValueError):
            OF.gaussian_filter(c, (1,), s, 3)


# ----------------------------------------------------------------------------- BPE
def test_bpe_two_cell_column():
    """test/test_ape_diagnostics.jl:85-107."""
    g = OGrid(2, z=(-1, 0), topology=("F", "F", "B"))
    b = OField(g, "ccc")
    b.interior[0, 0, :] = [1.0, -1.0]
    r = OB.reference_height(b)
    np.testing.assert_allclose(r["zstar"].ravel(), [-0.25, -0.75], rtol=1e-14)
    Eb = K.integral(-b.interior * r["zstar"], g, "ccc")
    assert Eb == pytest.approx(-0.25)
    b.interior[0, 0, :] = [-1.0, 1.0]
    r = OB.reference_height(b)
    np.testing.assert_allclose(r["zstar"].ravel(), [-0.75, -0.25], rtol=1e-14)


def _profile(b, r):
    h = r["zstar"].ravel(order="F")
    bb = b.interior.ravel(order="F")
    o = np.argsort(h, kind="stable")
    return h[o], bb[o]


def test_bpe_known_tanh_profile_and_ties():
    """test/test_ape_diagnostics.jl:405-461."""
    Nx, Ny, Nz = 4, 3, 5
    Nc = Nx * Ny * Nz
    Lx, Ly, Lz = 2, 3, 1
    g = OGrid((Nx, Ny, Nz), x=(0, Lx), y=(0, Ly), z=(-Lz, 0))
    B0, h, zmid = 0.1, 0.2, -Lz / 2
    dzs = Lz / Nc
    dV = Lx * Ly * Lz / Nc
    zs = -Lz + (np.arange(1, Nc + 1) - 0.5) * dzs
    bs = B0 * np.tanh((zs - zmid) / h)
    scr = np.zeros(Nc)
    for m in range(Nc):
        scr[(7 * m) % Nc] = bs[m]
    b = OField(g, "ccc")
    b.interior[...] = scr.reshape((Nx, Ny, Nz), order="F")
    Eb_exp = np.sum(-bs * zs * dV)
    for method in ("ThreeDimensionalSort", "HeavisideIntegral"):
        r = OB.reference_height(b, method)
        hh, bb = _profile(b, r)
        np.testing.assert_allclose(hh, zs, rtol=1e-12)
        np.testing.assert_allclose(bb, bs, rtol=1e-12)
        assert K.integral(-b.interior * r["zstar"], g, "ccc") == pytest.approx(Eb_exp, rel=1e-12)
        # sorting moves no volume (test/test_ape_diagnostics.jl:71)
        zc = np.broadcast_to(g.node(2, "c", g.indices("ccc")[2]), r["zstar"].shape)
        assert K.average(r["zstar"], g, "ccc") == pytest.approx(K.average(zc, g, "ccc"), abs=1e-8)
        # local APE eₐ = Ψδ - b (z - z✶) >= 0 (src/AvailablePotentialEnergyEquation.jl:42)
        assert (r["psi_delta"] - b.interior * (zc - r["zstar"])).min() > -1e-12
    levels = 12
    tied = B0 * np.round(np.tanh((zs - zmid) / h) * levels) / levels
    for m in range(Nc):
        scr[(7 * m) % Nc] = tied[m]
    b.interior[...] = scr.reshape((Nx, Ny, Nz), order="F")
    tied_depth = max(np.sum(tied == v) for v in np.unique(tied)) * dzs
    rs = [OB.reference_height(b, m) for m in ("ThreeDimensionalSort", "HeavisideIntegral")]
    (h1, b1), (h2, b2) = (_profile(b, r) for r in rs)
    np.testing.assert_allclose(b2, b1, rtol=0, atol=0)
    assert np.max(np.abs(h2 - h1)) <= tied_depth
    E = [K.integral(-b.interior * r["zstar"], g, "ccc") for r in rs]
    assert E[1] == pytest.approx(E[0], rel=1e-12)


def test_bpe_heaviside_is_identity_on_stable_stratification_and_runs_stretched():
    """test/test_ape_diagnostics.jl:197-218, 368-390."""
    g = regular_grid()
    b = OField(g, "ccc").set(lambda x, y, z: 3 * z)
    r = OB.reference_height(b, "HeavisideIntegral")
    zc = np.broadcast_to(g.node(2, "c", g.indices("ccc")[2]), r["zstar"].shape)
    np.testing.assert_allclose(r["zstar"], zc, rtol=1e-12)
    r3 = OB.reference_height(b, "ThreeDimensionalSort")
    assert np.max(np.abs(r3["zstar"] - zc)) <= (1 / N) / 2 + 1e-14
    gs = stretched_grid()
    bs = OField(gs, "ccc")
    bs.interior[...] = np.random.default_rng(3).standard_normal(bs.interior.shape)
    with pytest.raises(ValueError, match="uniform cell volumes"):
        OB.reference_height(bs, "ThreeDimensionalSort")
    rr = OB.reference_height(bs, "HeavisideIntegral")
    zb, zt = gs.faces[2][1], gs.faces[2][N + 1]
    assert np.all(np.isfinite(rr["zstar"])) and rr["zstar"].min() >= zb and rr["zstar"].max() <= zt
    zc = np.broadcast_to(gs.node(2, "c", gs.indices("ccc")[2]), rr["zstar"].shape)
    assert K.average(rr["zstar"], gs, "ccc") == pytest.approx(K.average(zc, gs, "ccc"), abs=1e-8)


# ----------------------------------------------------------------------------- APE (scope table 8f, row 1)
def _as_field(grid, values):
    f = OField(grid, "ccc")
    f.interior[...] = values
    f.fill_halo()
    return f


def test_ape_two_cell_column_known_answers():
    """test/test_ape_diagnostics.jl:85-107: ∫E_p = +0.25, ∫E_b = -0.25, ∫E_a = +0.50 for the unstable column; the
    stable column holds no available potential energy."""
    g = OGrid(2, z=(-1, 0), topology=("F", "F", "B"))
    b = OField(g, "ccc")
    b.interior[0, 0, :] = [1.0, -1.0]
    r = OB.reference_height(b)
    z = g.node(2, "c", np.arange(1, 3))
    assert K.integral(-b.interior * z.reshape(1, 1, 2), g, "ccc") == pytest.approx(0.25)
    psi, zs = _as_field(g, r["psi_delta"]), _as_field(g, r["zstar"])
    Ea = K.evaluate(K.local_ape_ccc, g, "ccc", psi, b, zs)
    assert K.integral(Ea, g, "ccc") == pytest.approx(0.50)
    b.interior[0, 0, :] = [-1.0, 1.0]
    r = OB.reference_height(b)
    Ea = K.evaluate(K.local_ape_ccc, g, "ccc", _as_field(g, r["psi_delta"]), b, _as_field(g, r["zstar"]))
    assert abs(K.integral(Ea, g, "ccc")) <= 1e-14


@pytest.mark.parametrize("method", ["ThreeDimensionalSort", "HeavisideIntegral"])
def test_ape_vanishes_when_sorted_and_is_nonnegative(method):
    """test/test_ape_diagnostics.jl:108-122 (b = 3z is its own sorted state) and :61-76 (e_a >= 0, mean z✶ = mean z)."""
    g = OGrid((4, 3, 8), x=(0, 1), y=(0, 1), z=(-1, 0), topology=("P", "P", "B"))
    b = OField(g, "ccc")
    b.set(lambda x, y, z: 3 * z + 0 * x)
    b.fill_halo()
    r = OB.reference_height(b, method)
    Ea = K.evaluate(K.local_ape_ccc, g, "ccc", _as_field(g, r["psi_delta"]), b, _as_field(g, r["zstar"]))
    Ep = K.integral(-b.interior * g.node(2, "c", np.arange(1, 9)).reshape(1, 1, 8), g, "ccc")
    assert abs(K.integral(Ea, g, "ccc")) <= np.sqrt(np.finfo(float).eps) * abs(Ep)
    b.interior[...] = np.random.default_rng(5).standard_normal(b.interior.shape)
    b.fill_halo()
    r = OB.reference_height(b, method)
    zs = _as_field(g, r["zstar"])
    Ea = K.evaluate(K.local_ape_ccc, g, "ccc", _as_field(g, r["psi_delta"]), b, zs)
    assert Ea.min() > -np.sqrt(np.finfo(float).eps) * np.abs(b.interior).max()
    ups = K.evaluate(K.upsilon_ccc, g, "ccc", zs)                      # :541-563: Υ = z✶ - z, zero volume mean
    np.testing.assert_allclose(ups, r["zstar"] - g.node(2, "c", np.arange(1, 9)).reshape(1, 1, 8), rtol=0, atol=1e-15)
    assert abs(K.average(ups, g, "ccc")) <= np.sqrt(np.finfo(float).eps)


def test_ape_dissipation_matches_tracer_variance_discretization():
    """test/test_ape_diagnostics.jl:565-587: handed b in place of Υ, ε_a = χ / 2 to roundoff; :592-608: it vanishes for a
    sorted state."""
    from oracle.kernels import Closure
    g = OGrid((5, 4, 6), x=(0, 1), y=(0, 1), z=(-1, 0), topology=("P", "P", "B"))
    b = OField(g, "ccc")
    b.interior[...] = np.random.default_rng(7).standard_normal(b.interior.shape)
    b.fill_halo()
    cl = Closure(nu=1e-6, kappa=1e-3)
    eps_a = K.evaluate(K.ape_dissipation_rate_ccc, g, "ccc", b, cl, None, None, b)
    chi = K.evaluate(K.tracer_variance_dissipation_rate_ccc, g, "ccc", cl, None, None, b)
    np.testing.assert_allclose(eps_a, chi / 2, rtol=1e-13, atol=1e-18)
    b.set(lambda x, y, z: 3 * z + 0 * x)
    b.fill_halo()
    r = OB.reference_height(b, "HeavisideIntegral")
    ups = _as_field(g, r["zstar"] - g.node(2, "c", np.arange(1, 7)).reshape(1, 1, 6))
    chi = K.evaluate(K.tracer_variance_dissipation_rate_ccc, g, "ccc", cl, None, None, b)
    eps_a = K.evaluate(K.ape_dissipation_rate_ccc, g, "ccc", ups, cl, None, None, b)
    assert chi.max() > 0 and np.abs(eps_a).max() < np.sqrt(np.finfo(float).eps) * chi.max() / 2


def _stride_scrambled(values, shape):
    N = values.size
    scr = np.zeros(N)
    for m in range(N):
        scr[(7 * m) % N] = values[m]                     # a bijection when gcd(7, N) = 1
    return scr.reshape(shape, order="F")


def test_vertical_sort_recovers_the_known_profile():
    """test/test_ape_diagnostics.jl:405-441: a scrambled tanh profile: the column's heights, buoyancies and ∫E_b."""
    Nx, Ny, Nz = 4, 3, 5
    N, Lz = Nx * Ny * Nz, 1.0
    g = OGrid((Nx, Ny, Nz), x=(0, 2), y=(0, 3), z=(-Lz, 0), topology=("P", "P", "B"))
    zs = -Lz + (np.arange(1, N + 1) - 0.5) * Lz / N
    bs = 0.1 * np.tanh((zs + Lz / 2) / 0.2)
    b = OField(g, "ccc")
    b.interior[...] = _stride_scrambled(bs, (Nx, Ny, Nz))
    col = OB.vertical_sort(b)
    np.testing.assert_allclose(col["zstar"], zs, rtol=1e-13)
    np.testing.assert_allclose(col["b_sorted"], bs, rtol=0, atol=0)
    assert np.sum(-col["b_sorted"] * col["zstar"] * (2 * 3 * Lz / N)) == pytest.approx(np.sum(-bs * zs * (2 * 3 * Lz / N)), rel=1e-13)
    assert np.all(np.diff(col["b_sorted"]) >= 0) and np.all(np.diff(col["zstar"]) > 0)          # :291-293
    # the column integrates like the model grid (:295)
    assert np.sum(col["b_sorted"]) == pytest.approx(np.sum(b.interior), rel=1e-13)


def test_profile_lookup_matches_the_ranked_sort_and_validates_profiles():
    """test/test_ape_diagnostics.jl:226-275: on a tie-free field ProfileLookup -- from the field's own sort or a borrowed
    column / bare pair -- reproduces ThreeDimensionalSort exactly; malformed profiles are rejected with the reference's
    messages."""
    Nx, Ny, Nz = 3, 2, 4
    N = Nx * Ny * Nz
    g = OGrid((Nx, Ny, Nz), x=(0, 1), y=(0, 1), z=(-1, 0), topology=("P", "P", "B"))
    b = OField(g, "ccc")
    b.interior[...] = _stride_scrambled(-0.5 + np.arange(N) / (N - 1), (Nx, Ny, Nz))
    ranked = OB.reference_height(b, "ThreeDimensionalSort")
    col = OB.vertical_sort(b)
    got = OB.profile_lookup(b, col["b_sorted"], col["zstar"])
    np.testing.assert_allclose(got["zstar"], ranked["zstar"], rtol=1e-13)
    np.testing.assert_allclose(got["psi_delta"], ranked["psi_delta"], rtol=1e-12, atol=1e-15)
    bs, zp = col["b_sorted"], col["zstar"]
    for msg, bad in (("buoyancy and height have different", (bs[:-1], zp)), ("ordered from the densest fluid up", (bs[::-1], zp)),
                     ("heights paired with", (bs, zp[::-1]))):
        with pytest.raises(ValueError, match=msg):
            OB.profile_lookup(b, *bad)
