"""GPU parity tests of the fused stencil sweep (K1), through the C ABI (ocb_plan_create / execute), against the
NumPy oracle on identical seeded inputs.  Float64: rtol 1e-12 relative to the field magnitude; Float32: 1e-5
relative L2 (BASELINE.json north_star)."""
import numpy as np
import pytest
import torch

from helpers import State, run_ops, assert_close, K, ocb
from cases import build_cases, integral_scale, check_case

pytestmark = pytest.mark.gpu


def _check(st, cases, vals, ints, og, rtol=1e-12):
    for n, ((name, op, want, loc), got) in enumerate(zip(cases, vals)):
        finite = np.isfinite(want)
        check_case(st, name, got, want, rtol)
        if ints is not None and finite.all() and not name.startswith("RichardsonNumber"):
            assert abs(ints[n] - K.integral(want, og, loc)) <= rtol * integral_scale(og, want, loc), name


@pytest.mark.parametrize("stretched", [False, True], ids=["regular", "stretched"])
@pytest.mark.parametrize("closure", ["scalar", "smagorinsky_like", "tuple"])
def test_every_diagnostic_fused_matches_oracle(stretched, closure):
    st = State(size=(40, 24, 18), stretched=stretched, device="cuda")
    cases = build_cases(st, closure, mean="profile")
    vals, ints = run_ops([c[1] for c in cases], st.pg, integrals=True, backend="cuda", fused=True)
    _check(st, cases, vals, ints, st.og)


@pytest.mark.parametrize("mean", ["number", "none"])
def test_unfused_and_operand_kinds(mean):
    st = State(size=(33, 17, 9), device="cuda")       # ragged sizes: partial blocks in x and y
    cases = build_cases(st, "scalar", mean=mean)
    vals, ints = run_ops([c[1] for c in cases], st.pg, integrals=True, backend="cuda", fused=False)
    _check(st, cases, vals, ints, st.og)


def test_ke_forcing_functors_known_answer():
    """test/test_kinetic_energy_equation.jl:105-129 through the package: Forcing(func, field_dependencies) functors are
    materialised at the velocity points (halos included) and uᵢFᵤᵢ = @at ccc (-u² - v² - w²) to rtol 1e-12; the forcing
    fields follow the velocities when these change (computed operands are refreshed before the sweep)."""
    st = State(size=(40, 24, 18), device="cuda")
    damp = lambda x, y, z, t, q: -q
    m = st.model(closure=st.closures()["scalar"][1])
    m.forcing.u, m.forcing.v, m.forcing.w = (ocb.Forcing(damp, field_dependencies=n) for n in ("u", "v", "w"))
    forc = ocb.Field(ocb.KineticEnergyEquation.KineticEnergyForcing(m))
    ke = ocb.Field(ocb.KineticEnergyEquation.KineticEnergy(m))
    forc.compute(); ke.compute()
    torch.cuda.synchronize()
    np.testing.assert_allclose(forc.interior_ijk(), -2.0 * ke.interior_ijk(), rtol=1e-12, atol=1e-300)
    og, f = st.og, st.ofields()
    sq = lambda i, j, k, gg, q: -q[i, j, k] ** 2
    truth = (K.evaluate(K.ix_caa, og, "ccc", sq, st.ou) + K.evaluate(K.iy_aca, og, "ccc", sq, st.ov) + K.evaluate(K.iz_aac, og, "ccc", sq, st.ow))
    np.testing.assert_allclose(forc.interior_ijk(), truth, rtol=1e-12, atol=1e-300)
    m.velocities.u.data.mul_(3.0)
    forc.compute(time=1.0); ke.compute(time=1.0)
    torch.cuda.synchronize()
    np.testing.assert_allclose(forc.interior_ijk(), -2.0 * ke.interior_ijk(), rtol=1e-12, atol=1e-300)
    # a model without forcing: identically zero; hydrostatic models are refused as in the reference
    none = ocb.Field(ocb.KineticEnergyEquation.KineticEnergyForcing(st.model(closure=st.closures()["scalar"][1]))).compute()
    assert float(none.interior.abs().max()) == 0.0


def test_config1_64cube_dissipation_and_richardson():
    """BASELINE.json configs[0]: 64³ Float64, ScalarDiffusivity: DissipationRate and RichardsonNumber."""
    st = State(size=(64, 64, 64), device="cuda")
    oc, pc = st.closures()["scalar"]
    m = st.model(closure=pc)
    ops = [ocb.KineticEnergyEquation.DissipationRate(m), ocb.FlowDiagnostics.RichardsonNumber(m)]
    vals, _ = run_ops(ops, st.pg, backend="cuda")
    want_eps = K.evaluate(K.viscous_dissipation_rate_ccc, st.og, "ccc", None, st.ofields(), {"closure": oc})
    want_ri = K.evaluate(K.richardson_number_ccf, st.og, "ccf", st.ou, st.ov, st.ow, st.ob, (0, 0, 1))
    assert vals[0].shape == (64, 64, 64) and vals[1].shape == (64, 64, 65)
    np.testing.assert_allclose(vals[0], want_eps, rtol=1e-12, atol=0)      # ε is a sum of squares: true elementwise rtol
    check_case(st, "RichardsonNumber", vals[1], want_ri)


def test_float32_stretched_les_closures():
    """configs[3] at test size: stretched-z Float32 with eddy-viscosity closures: ε and |S| to 1e-5 relative L2."""
    st = State(size=(48, 40, 24), stretched=True, dtype=np.float32, device="cuda")
    for closure in ("smagorinsky_like", "tuple"):
        oc, pc = st.closures()[closure]
        m = st.model(closure=pc)
        ops = [ocb.KineticEnergyEquation.DissipationRate(m), ocb.FlowDiagnostics.StrainRateTensorModulus(m)]
        vals, _ = run_ops(ops, st.pg, backend="cuda")
        want = [K.evaluate(K.viscous_dissipation_rate_ccc, st.og, "ccc", None, st.ofields(), {"closure": oc}),
                K.evaluate(K.strain_rate_tensor_modulus_ccc, st.og, "ccc", st.ou, st.ov, st.ow)]
        for g, w_ in zip(vals, want):
            num = np.linalg.norm(g.astype(np.float64) - w_.astype(np.float64))
            assert num <= 1e-5 * np.linalg.norm(w_.astype(np.float64))


def test_windowed_destination_matches_full():
    """Field(op; indices=...) writes only its window (reference regression: test/test_spatial_filters.jl:210-222)."""
    st = State(size=(20, 16, 12), device="cuda")
    m = st.model(closure=st.closures()["scalar"][1])
    op = ocb.KineticEnergyEquation.DissipationRate(m)
    full = ocb.Field(op).compute()
    win = ocb.Field(op, indices=((3, 11), (2, 9), (5, 5)))
    win.data.fill_(-7.0)
    win.compute()
    torch.cuda.synchronize()
    a = full.interior_ijk()[2:11, 1:9, 4:5]
    np.testing.assert_allclose(win.interior_ijk(), a, rtol=1e-13, atol=0)   # full: pipelined kernel, window: generic kernel
    assert float(win.data[0, 0, 0]) == -7.0      # nothing outside the window's interior was touched


def test_computed_field_halos_are_filled():
    st = State(size=(12, 10, 8), device="cuda")
    m = st.model(closure=st.closures()["scalar"][1])
    f = ocb.Field(ocb.KineticEnergyEquation.KineticEnergy(m)).compute()
    torch.cuda.synchronize()
    d = f.data.cpu().numpy()       # [k, j, i]
    H = 3
    np.testing.assert_array_equal(d[H:-H, H:-H, 0:H], d[H:-H, H:-H, -2 * H:-H])      # periodic x
    np.testing.assert_array_equal(d[H - 1, H:-H, H:-H], d[H, H:-H, H:-H])           # no-flux bottom


def test_integrals_are_run_to_run_deterministic_and_fusion_invariant():
    st = State(size=(64, 48, 32), device="cuda")
    m = st.model(closure=st.closures()["scalar"][1])
    KE = ocb.KineticEnergyEquation
    ops = [KE.KineticEnergy(m), KE.KineticEnergyAdvection(m), KE.DissipationRate(m)]
    grp = ocb.FusedDiagnostics(*[ocb.Integral(o) for o in ops])
    first = grp.compute().integrals()
    for _ in range(3):
        assert grp.compute().integrals() == first
    single = [ocb.Field(ocb.Integral(o)).compute().value() for o in ops]   # other kernel variants: equal to rounding
    for a, b, o in zip(single, first, ops):
        scale = float(ocb.Field(o).compute().interior.abs().sum()) / (64 * 48 * 32)
        assert abs(a - b) <= 1e-13 * scale


def test_average_is_integral_over_volume():
    st = State(size=(16, 16, 8), stretched=True, device="cuda")
    m = st.model(closure=st.closures()["scalar"][1])
    op = ocb.KineticEnergyEquation.KineticEnergy(m)
    avg = ocb.Field(ocb.Average(op)).compute().value()
    want = K.average(K.evaluate(K.kinetic_energy_ccc, st.og, "ccc", st.ou, st.ov, st.ow), st.og, "ccc")
    assert abs(avg - want) <= 1e-12 * abs(want)


def test_status_gated_recompute():
    """compute_at! semantics: a fused group recomputes only when the time changes (BPE delegation pattern)."""
    st = State(size=(12, 12, 6), device="cuda")
    m = st.model(closure=st.closures()["scalar"][1])
    f = ocb.Field(ocb.KineticEnergyEquation.KineticEnergy(m))
    grp = ocb.FusedDiagnostics(f)
    grp.compute(time=1.0)
    before = f.interior_ijk().copy()
    st.u.data.mul_(2.0)
    grp.compute(time=1.0)                      # same time: no-op
    np.testing.assert_array_equal(f.interior_ijk(), before)
    grp.compute(time=2.0)                      # new time: recomputed from the evolved state
    assert not np.array_equal(f.interior_ijk(), before)
    # Oceananigans recomputes whenever time == 0 (compute_at!: `time == zero(time) || time != status.time`)
    grp.compute(time=0.0)
    at_zero = f.interior_ijk().copy()
    st.u.data.mul_(2.0)
    grp.compute(time=0.0)
    assert not np.array_equal(f.interior_ijk(), at_zero)
    ocb.compute_at(f, 0.0)


def test_error_behaviour():
    st = State(size=(8, 8, 4), device="cuda")
    m = st.model(closure=st.closures()["scalar"][1])
    with pytest.raises(ValueError):            # ArgumentError in the reference (src/Oceanostics.jl:114-116)
        ocb.KineticEnergyEquation.DissipationRate(m, location=("f", "c", "c"))
    with pytest.raises(ValueError):
        ocb.FlowDiagnostics.StrainRateTensor(m, dims=(1, 1))
    with pytest.raises(ValueError):
        ocb.Closure(isotropic=False, name="HorizontalScalarDiffusivity")
    bad = ocb.Field(ocb.KineticEnergyEquation.KineticEnergy(m), indices=((1, 9), None, None))   # window past Nx
    with pytest.raises(ValueError):
        bad.compute()
    # operands must carry the halo the stencils reach into: a one-cell halo serves the kinetic energy, not the advection term
    thin = State(size=(8, 8, 4), device="cuda", halo=(1, 1, 1))
    mt = thin.model(closure=thin.closures()["scalar"][1])
    ocb.Field(ocb.KineticEnergyEquation.KineticEnergy(mt)).compute()
    with pytest.raises(ValueError, match="halo of 1"):
        ocb.Field(ocb.KineticEnergyEquation.KineticEnergyAdvection(mt)).compute()
