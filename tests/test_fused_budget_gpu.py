"""GPU parity tests of the register-pipelined budget kernel (csrc/ocb_sweep_fused.cu): the fused
KE + tracer-variance budget of BASELINE.json configs[2] against the NumPy oracle (rtol 1e-12), against the generic
kernel at a size the oracle cannot reach, and through size-independent identities at larger sizes."""
import os

import numpy as np
import pytest
import torch

from helpers import State, assert_close, K, ocb
from cases import integral_scale

pytestmark = pytest.mark.gpu


def budget(st, closure="scalar"):
    oc, pc = st.closures()[closure]
    m = st.model(closure=pc)
    KE, TV = ocb.KineticEnergyEquation, ocb.TracerVarianceEquation
    ops = [KE.KineticEnergy(m), KE.KineticEnergyAdvection(m), KE.KineticEnergyPressureRedistribution(m),
           KE.PotentialEnergyConversion(m), KE.DissipationRate(m), TV.TracerVarianceDissipationRate(m, "b"),
           TV.TracerVarianceDiffusion(m, "b")]
    return ops, oc


def oracle_budget(st, oc):
    og, f = st.og, st.ofields()
    ev = K.evaluate
    return [ev(K.kinetic_energy_ccc, og, "ccc", st.ou, st.ov, st.ow), ev(K.ke_advection_ccc, og, "ccc", f),
            ev(K.ke_pressure_ccc, og, "ccc", f, st.op), ev(K.ke_buoyancy_ccc, og, "ccc", f, (0, 0, 1), st.ob),
            ev(K.viscous_dissipation_rate_ccc, og, "ccc", None, f, {"closure": oc}),
            ev(K.tracer_variance_dissipation_rate_ccc, og, "ccc", oc, None, None, st.ob),
            ev(K.c_div_q_c, og, "ccc", oc, None, None, st.ob)]


def run(ops, fields=True, integrals=True, disable_fused=False):
    if disable_fused:
        os.environ["OCB_DISABLE_FUSED"] = "1"
    try:
        members = ([ocb.Field(o) for o in ops] if fields else []) + ([ocb.Integral(o) for o in ops] if integrals else [])
        grp = ocb.FusedDiagnostics(*members)
    finally:
        os.environ.pop("OCB_DISABLE_FUSED", None)
    grp.compute()
    torch.cuda.synchronize()
    vals = [m.interior_ijk() for m in grp.members if isinstance(m, ocb.ComputedField)]
    return grp, vals, grp.integrals()


@pytest.mark.parametrize("size", [(64, 64, 64), (70, 45, 37), (32, 11, 5), (62, 23, 3), (33, 12, 3)], ids=str)
def test_fused_budget_matches_oracle(size):
    st = State(size=size, device="cuda")
    ops, oc = budget(st)
    grp, vals, ints = run(ops)
    # TMA needs 16-byte row pitches: an even Nx takes the pipelined kernel, an odd one the generic kernel
    assert grp.plan.variant == (1 if size[0] % 2 == 0 else 0)
    want = oracle_budget(st, oc)
    for op, g, w_, integ in zip(ops, vals, want, ints):
        assert_close(g, w_, 1e-12, op.name)
        assert abs(integ - K.integral(w_, st.og, "ccc")) <= 1e-12 * integral_scale(st.og, w_, "ccc"), op.name


@pytest.mark.parametrize("size", [(64, 48, 40), (34, 22, 9), (32, 11, 4)], ids=str)
@pytest.mark.parametrize("topology", [("P", "P", "B"), ("P", "B", "B")], ids=lambda t: "".join(t))
def test_fused_budget_stretched_z_matches_oracle(size, topology):
    """The per-plane-metric variant of the budget kernel (stretched z, the reference's tanh-like test grid of
    test/test_utils.jl:17-24): all seven terms as fields and as integrals against the oracle's area / volume weighted forms,
    rtol 1e-12 -- dz_f of the face in the strain edges and tracer gradients, dz_c of the cell in the divergences and the integral
    weights, dz_c-weighted advecting transports in the w equation."""
    st = State(size=size, stretched=True, device="cuda", topology=topology)
    ops, oc = budget(st)
    grp, vals, ints = run(ops)
    assert grp.plan.variant == 1
    want = oracle_budget(st, oc)
    for op, g, w_, integ in zip(ops, vals, want, ints):
        assert_close(g, w_, 1e-12, op.name)
        assert abs(integ - K.integral(w_, st.og, "ccc")) <= 1e-12 * integral_scale(st.og, w_, "ccc"), op.name
    # integrals only (no destinations), and a subset
    grp, _, ints2 = run(ops, fields=False)
    assert grp.plan.variant == 1
    for a, b_, w_ in zip(ints, ints2, want):
        assert abs(a - b_) <= 1e-12 * integral_scale(st.og, w_, "ccc")          # (the class-sum form: other association)
    for subset in ([1, 4], [0, 1, 2, 4], [1, 3, 5, 6]):          # velocity-only, + pressure, + tracer variants
        grp, vals, _ = run([ops[q] for q in subset], integrals=False)
        assert grp.plan.variant == 1
        for q, g in zip(subset, vals):
            assert_close(g, want[q], 1e-12, ops[q].name)


def test_fused_budget_stretched_z_matches_generic_kernel_at_size():
    st = State(size=(128, 96, 64), stretched=True, device="cuda")
    ops, _ = budget(st)
    grp, vals, ints = run(ops)
    assert grp.plan.variant == 1
    grp0, vals0, ints0 = run(ops, disable_fused=True)
    assert grp0.plan.variant in (0, 4)
    for op, a, b_, ia, ib in zip(ops, vals, vals0, ints, ints0):
        assert_close(a, b_, 1e-12, op.name)
        assert abs(ia - ib) <= 1e-12 * max(abs(ia), abs(ib), float(np.abs(b_).sum()) * 1e-6), op.name


def test_fused_integrals_only_and_subsets():
    st = State(size=(48, 40, 20), device="cuda")
    ops, oc = budget(st)
    want = oracle_budget(st, oc)
    grp, _, ints = run(ops, fields=False)
    assert grp.plan.variant == 1 and grp.plan.algorithmic_bytes == 5 * 8 * 48 * 40 * 20
    for op, w_, integ in zip(ops, want, ints):
        assert abs(integ - K.integral(w_, st.og, "ccc")) <= 1e-12 * integral_scale(st.og, w_, "ccc"), op.name
    # dissipation alone (configs[0]) takes the dissipation-only variant
    grp, vals, _ = run([ops[4]], integrals=False)
    assert grp.plan.variant == 1
    np.testing.assert_allclose(vals[0], want[4], rtol=1e-12, atol=0)
    # a subset with no pressure operand
    grp, vals, _ = run([ops[0], ops[1], ops[5]], integrals=False)
    assert grp.plan.variant == 1
    for g, w_ in zip(vals, [want[0], want[1], want[5]]):
        assert_close(g, w_, 1e-12)


def test_fused_agrees_with_generic_kernel_at_scale():
    """Two independent CUDA formulations (shared-intermediate pipeline vs per-cell evaluators) at 200x150x96."""
    st = State(size=(200, 150, 96), device="cuda")
    ops, _ = budget(st)
    g1, v1, i1 = run(ops)
    g0, v0, i0 = run(ops, disable_fused=True)
    assert g1.plan.variant == 1 and g0.plan.variant == 0
    for op, a, b in zip(ops, v1, v0):
        assert_close(a, b, 1e-12, op.name)
    for a, b, v in zip(i1, i0, v0):
        assert abs(a - b) <= 1e-12 * float(np.sum(np.abs(v))) * (1.0 / (200 * 150 * 96))


def test_fused_is_deterministic_and_identities_hold():
    """Size-independent properties at 256x256x128: run-to-run bit equality; <chi> = <2c div F> (the reference's own
    identity, test/test_tracer_variance_diagnostics.jl:31-34); dissipation is non-negative."""
    st = State(size=(256, 256, 128), device="cuda")
    ops, _ = budget(st)
    grp, _, ints = run(ops, fields=False)
    for _ in range(3):
        grp.compute()
        assert grp.integrals() == ints
    chi, cdf = ints[5], ints[6]
    assert abs(chi - cdf) <= 1e-10 * abs(chi)
    f = ocb.Field(ops[4]).compute()
    assert float(f.interior.min()) >= 0.0


def test_budget_kernel_envelope():
    st = State(size=(40, 24, 18), stretched=True, device="cuda")
    ops, _ = budget(st)
    grp, _, _ = run(ops, fields=False)
    assert grp.plan.variant == 1          # stretched z, constant coefficients: the per-plane-metric variant of the budget kernel
    st = State(size=(40, 24, 18), device="cuda")
    oc, pc = st.closures()["smagorinsky_like"]
    m = st.model(closure=pc)
    grp = ocb.FusedDiagnostics(ocb.Integral(ocb.KineticEnergyEquation.DissipationRate(m)))
    assert grp.plan.variant == 3          # eddy-viscosity field, dissipation alone (no b, p operands): velocity-gradient kernel


@pytest.mark.parametrize("stretched", [False, True], ids=["regular", "stretched"])
def test_les_closure_budget_splits_over_the_pipelined_kernels(stretched):
    """Eddy-viscosity closures (a nu_e / kappa_e field): the closure-independent terms (K, ADV, PR, w b) stay on the budget kernel, the
    dissipation goes to the velocity-gradient kernel, the tracer-variance terms to the generic one -- one plan, three launches,
    every term against the oracle."""
    st = State(size=(48, 40, 20), stretched=stretched, device="cuda")
    ops, oc = budget(st, closure="smagorinsky_like")
    grp, vals, ints = run(ops)
    assert grp.plan.variant == 6
    want = oracle_budget(st, oc)
    for op, g, w_, integ in zip(ops, vals, want, ints):
        assert_close(g, w_, 1e-12, op.name)
        assert abs(integ - K.integral(w_, st.og, "ccc")) <= 1e-12 * integral_scale(st.og, w_, "ccc"), op.name
    grp, _, ints2 = run(ops, fields=False)     # integrals only: ONE launch of the class-sum kernel, nu_e inside the dissipation's
    assert grp.plan.variant == 1               # items and kappa_e = nu_e / Pr on the tracer gradients' faces
    for op, w_, integ in zip(ops, want, ints2):
        assert abs(integ - K.integral(w_, st.og, "ccc")) <= 1e-12 * integral_scale(st.og, w_, "ccc"), op.name
    grp, vals, _ = run(ops[:5], integrals=False)                 # the KE budget alone: no generic launch
    assert grp.plan.variant == 5
    for op, g, w_ in zip(ops[:5], vals, want[:5]):
        assert_close(g, w_, 1e-12, op.name)


@pytest.mark.parametrize("stretched", [False, True], ids=["regular", "stretched"])
@pytest.mark.parametrize("closure", ["smagorinsky_like", "tuple"])
def test_eddy_viscosity_dissipation_as_class_sums(stretched, closure):
    """The KE + tracer-variance budget integrals under SmagorinskyLilly- / AMD-like closures (a nu_e field, alone or in a tuple
    with constant coefficients; kappa_e = nu_e / Pr) on the class-sum kernel: nu at ccc for the diagonal strains, its four-point
    means at ffc / fcf / cff for the edges, kappa on each tracer-gradient face, all inside the sums.  Whole domain and random
    y / z windows against the oracle's per-cell values times their volumes and against the split plan (budget kernel +
    velocity-gradient kernel + generic kernel)."""
    st = State(size=(64, 40, 30), stretched=stretched, device="cuda")
    ops, oc = budget(st, closure=closure)
    want = oracle_budget(st, oc)
    ii, jj, kk = st.og.indices("ccc")
    vol = np.broadcast_to(st.og.V("ccc", ii, jj, kk), want[0].shape)
    Nx, Ny, Nz = 64, 40, 30
    rng = np.random.default_rng(7)
    windows = [None, (1, Ny, 1, 1), (1, 1, 1, Nz), (2, Ny - 1, 2, Nz - 1), (Ny, Ny, Nz, Nz)]
    for _ in range(3):
        j = sorted(rng.integers(1, Ny + 1, 2)); k = sorted(rng.integers(1, Nz + 1, 2))
        windows.append((int(j[0]), int(j[1]), int(k[0]), int(k[1])))
    ke = ops                                                        # K, ADV, PR, w b, eps, chi, 2 c div(kappa grad c)
    os.environ["OCB_FUSED_KCHUNK"] = "16"
    try:
        for win in windows:
            idx = None if win is None else (None, (win[0], win[1]), (win[2], win[3]))
            jlo, jhi, klo, khi = win or (1, Ny, 1, Nz)
            grp = ocb.FusedDiagnostics(*[ocb.Integral(o, indices=idx) for o in ke]).compute()
            assert grp.plan.variant == 1 and grp.plan.n_launches == 2, (win, grp.plan.variant)
            os.environ["OCB_DISABLE_SUMS_NUF"] = "1"
            try:
                split = ocb.FusedDiagnostics(*[ocb.Integral(o, indices=idx) for o in ke]).compute()
            finally:
                os.environ.pop("OCB_DISABLE_SUMS_NUF", None)
            assert split.plan.variant == 6                          # budget kernel + velocity-gradient kernel + generic kernel
            for o, w_, a, b_ in zip(ke, want, grp.integrals(), split.integrals()):
                ww, vw = w_[:, jlo - 1:jhi, klo - 1:khi], vol[:, jlo - 1:jhi, klo - 1:khi]
                ref, scale = float(np.sum(ww * vw)), float(np.sum(np.abs(ww) * vw))
                assert abs(a - ref) <= 1e-12 * scale, (o.name, win, a, ref, scale)
                assert abs(a - b_) <= 1e-12 * scale, (o.name, win, a, b_, scale)
    finally:
        os.environ.pop("OCB_FUSED_KCHUNK", None)


@pytest.mark.parametrize("stretched", [False, True], ids=["regular", "stretched"])
@pytest.mark.parametrize("closure", ["smagorinsky_like", "tuple"])
def test_tke_budget_integrals_under_eddy_viscosity_closures(stretched, closure):
    """The boundary-layer LES case: the TKE budget about horizontal-mean profiles {shear production, dissipation of u - U,
    pressure redistribution, TKE} as integrals under an eddy-viscosity closure, regular or stretched z, in ONE launch of the
    class-sum kernel (nu_e a TMA-fed input, its means on the strain edges); against the oracle and the generic kernel."""
    from cases import build_cases
    st = State(size=(64, 44, 26), stretched=stretched, device="cuda")
    cases = {c[0]: c for c in build_cases(st, closure, mean="profile")}
    names = ["ShearProductionRate", "DissipationRate_perturbation", "KineticEnergyPressureRedistribution", "TurbulentKineticEnergy"]
    ops = [cases[nm][1] for nm in names]
    ii, jj, kk = st.og.indices("ccc")
    V = np.broadcast_to(st.og.V("ccc", ii, jj, kk), cases[names[0]][2].shape)
    Ny, Nz = 44, 26
    for win in (None, (3, 40, 2, 25), (1, 1, 1, Nz), (Ny, Ny, Nz, Nz), (10, 30, 13, 13)):
        idx = None if win is None else (None, (win[0], win[1]), (win[2], win[3]))
        jlo, jhi, klo, khi = win or (1, Ny, 1, Nz)
        grp = ocb.FusedDiagnostics(*[ocb.Integral(o, indices=idx) for o in ops]).compute()
        assert grp.plan.variant == 1 and grp.plan.n_launches == 2, (win, grp.plan.variant)
        os.environ["OCB_DISABLE_FUSED"] = "1"
        try:
            gen = ocb.FusedDiagnostics(*[ocb.Integral(o, indices=idx) for o in ops]).compute()
        finally:
            os.environ.pop("OCB_DISABLE_FUSED", None)
        for nm, a, b_ in zip(names, grp.integrals(), gen.integrals()):
            w_, v_ = cases[nm][2][:, jlo - 1:jhi, klo - 1:khi], V[:, jlo - 1:jhi, klo - 1:khi]
            ref, scale = float(np.sum(w_ * v_)), float(np.sum(np.abs(w_) * v_))
            assert abs(a - ref) <= 1e-12 * scale, (nm, win, a, ref, scale)
            assert abs(a - b_) <= 1e-12 * scale, (nm, win, a, b_, scale)


@pytest.mark.parametrize("stretched", [False, True], ids=["regular", "stretched"])
@pytest.mark.parametrize("closure", ["scalar", "smagorinsky_like"])
def test_float32_budget_integrals_on_the_class_sum_kernel(stretched, closure):
    """Float32 fields (BASELINE configs[3]'s precision): the seven budget integrals on the class-sum kernel's Float32-input form --
    TMA tiles as two boxes per field over the merged-row view (a 774-float row is not a 16-byte multiple), sums in Float64 --
    regular / stretched z, constant / eddy-viscosity closure, whole domain and windows.  Against the oracle's Float32 per-cell
    values times their volumes (1e-5 of the integrand scale, north_star's Float32 tolerance) and against the generic kernel."""
    st = State(size=(64, 40, 30), stretched=stretched, dtype=np.float32, device="cuda")
    ops, oc = budget(st, closure=closure)
    want = [w.astype(np.float64) for w in oracle_budget(st, oc)]
    ii, jj, kk = st.og.indices("ccc")
    vol = np.broadcast_to(np.asarray(st.og.V("ccc", ii, jj, kk), dtype=np.float64), want[0].shape)
    Ny, Nz = 40, 30
    for win in (None, (3, 38, 2, 29), (1, 1, 1, Nz), (Ny, Ny, Nz, Nz), (7, 21, 11, 11)):
        idx = None if win is None else (None, (win[0], win[1]), (win[2], win[3]))
        jlo, jhi, klo, khi = win or (1, Ny, 1, Nz)
        grp = ocb.FusedDiagnostics(*[ocb.Integral(o, indices=idx) for o in ops]).compute()
        assert grp.plan.variant == 1 and grp.plan.n_launches == 2, (win, grp.plan.variant)
        os.environ["OCB_DISABLE_SUMS_F32"] = "1"
        try:
            gen = ocb.FusedDiagnostics(*[ocb.Integral(o, indices=idx) for o in ops]).compute()
        finally:
            os.environ.pop("OCB_DISABLE_SUMS_F32", None)
        assert gen.plan.variant != 1
        for o, w_, a, b_ in zip(ops, want, grp.integrals(), gen.integrals()):
            ww, vw = w_[:, jlo - 1:jhi, klo - 1:khi], vol[:, jlo - 1:jhi, klo - 1:khi]
            ref, scale = float(np.sum(ww * vw)), float(np.sum(np.abs(ww) * vw))
            assert abs(a - ref) <= 1e-5 * scale, (o.name, win, a, ref, scale)
            assert abs(a - b_) <= 1e-5 * scale, (o.name, win, a, b_, scale)


@pytest.mark.parametrize("stretched", [False, True], ids=["regular", "stretched"])
@pytest.mark.parametrize("closure", ["scalar", "smagorinsky_like"])
def test_float32_tke_budget_integrals(stretched, closure):
    """Float32 fields and Float32 mean profiles: the TKE budget {shear production, dissipation of u - U, pressure redistribution,
    TKE} as integrals in one launch of the Float32-input class-sum kernel (the profiles are copied to Float64 by every execute);
    recomputed after the state changes; against the Float32 oracle and the generic kernel to 1e-5 of the integrand scale."""
    from cases import build_cases
    st = State(size=(64, 44, 26), stretched=stretched, dtype=np.float32, device="cuda")
    cases = {c[0]: c for c in build_cases(st, closure, mean="profile")}
    names = ["ShearProductionRate", "DissipationRate_perturbation", "KineticEnergyPressureRedistribution", "TurbulentKineticEnergy"]
    ops = [cases[nm][1] for nm in names]
    ii, jj, kk = st.og.indices("ccc")
    V = np.broadcast_to(np.asarray(st.og.V("ccc", ii, jj, kk), dtype=np.float64), cases[names[0]][2].shape)
    grp = ocb.FusedDiagnostics(*[ocb.Integral(o) for o in ops]).compute()
    assert grp.plan.variant == 1 and grp.plan.n_launches == 2
    os.environ["OCB_DISABLE_SUMS_F32"] = "1"
    try:
        gen = ocb.FusedDiagnostics(*[ocb.Integral(o) for o in ops]).compute()
    finally:
        os.environ.pop("OCB_DISABLE_SUMS_F32", None)
    assert gen.plan.variant != 1
    for nm, a, b_ in zip(names, grp.integrals(), gen.integrals()):
        w_ = cases[nm][2].astype(np.float64)
        ref, scale = float(np.sum(w_ * V)), float(np.sum(np.abs(w_) * V))
        assert abs(a - ref) <= 1e-5 * scale, (nm, a, ref, scale)
        assert abs(a - b_) <= 1e-5 * scale, (nm, a, b_, scale)
    # the profiles follow the state: scale u, recompute both ways
    st.u.data.mul_(1.5)
    a2 = grp.compute(time=1.0).integrals()
    b2 = gen.compute(time=1.0).integrals()
    for nm, a, b_, a0 in zip(names, a2, b2, grp.integrals()):
        scale = float(np.sum(np.abs(cases[nm][2].astype(np.float64)) * V))
        assert abs(a - b_) <= 2e-5 * 2.25 * scale, (nm, a, b_)


@pytest.mark.parametrize("stretched", [False, True], ids=["regular", "stretched"])
def test_float32_budget_fields_on_the_pipelined_kernel(stretched):
    """Float32 fields in and out: all seven budget terms as fields and integrals in one launch of the budget kernel's Float32 form
    (two-box TMA tiles, Float64 arithmetic, Float32 stores), regular and stretched z; 1e-5 relative L2 against the Float32 oracle
    (north_star's Float32 tolerance), integrals to 1e-5 of the integrand scale, and a windowed output."""
    st = State(size=(64, 40, 30), stretched=stretched, dtype=np.float32, device="cuda")
    ops, oc = budget(st)
    grp, vals, ints = run(ops)
    assert grp.plan.variant == 1
    want = [w.astype(np.float64) for w in oracle_budget(st, oc)]
    ii, jj, kk = st.og.indices("ccc")
    vol = np.broadcast_to(np.asarray(st.og.V("ccc", ii, jj, kk), dtype=np.float64), want[0].shape)
    for op, g, w_, integ in zip(ops, vals, want, ints):
        assert g.dtype == np.float32
        err = np.linalg.norm(g.astype(np.float64) - w_) / np.linalg.norm(w_)
        assert err <= 1e-5, (op.name, err)
        assert abs(integ - float(np.sum(w_ * vol))) <= 1e-5 * float(np.sum(np.abs(w_) * vol)), op.name
    win = ocb.FusedDiagnostics(*[ocb.Field(o, indices=(None, (5, 33), (4, 27))) for o in ops]).compute()
    torch.cuda.synchronize()
    assert win.plan.variant == 1
    for o, f, w in zip(ops, vals, win.members):
        np.testing.assert_array_equal(w.interior_ijk(), f[:, 4:33, 3:27])


@pytest.mark.parametrize("stretched", [False, True], ids=["regular", "stretched"])
def test_windowed_budget_fields_match_the_full_fields(stretched):
    """Field(op, indices=(:, j-window, k-window)) of all seven terms in one sweep of the pipelined kernel: the window holds
    exactly the full field's values (the z window of a stretched grid starts at a plane with other metrics than plane 1)."""
    st = State(size=(64, 40, 30), stretched=stretched, device="cuda")
    ops, _ = budget(st)
    _, full, _ = run(ops, integrals=False)
    win = ocb.FusedDiagnostics(*[ocb.Field(o, indices=(None, (5, 33), (4, 27))) for o in ops]).compute()
    torch.cuda.synchronize()
    assert win.plan.variant == 1
    for o, f, w in zip(ops, full, win.members):
        a, b = f[:, 4:33, 3:27], w.interior_ijk()
        assert b.shape == a.shape, o.name
        assert np.abs(a - b).max() <= 1e-13 * np.abs(a).max(), o.name


@pytest.mark.parametrize("stretched", [False, True], ids=["regular", "stretched"])
def test_y_and_z_windows_tile_the_domain(stretched):
    """Integral(op, indices=...) over y strips / z slabs (the pieces the overlapped multi-GPU step and the host-streamed
    path sweep separately) add up to the whole-domain integrals, all on the pipelined kernel (stretched z: its per-plane-metric
    variant, whose z slabs start and end at planes with different metrics)."""
    st = State(size=(64, 40, 30), stretched=stretched, device="cuda")
    ops, _ = budget(st)
    ref = ocb.FusedDiagnostics(*[ocb.Integral(o) for o in ops]).compute().integrals()
    for windows in ([(None, (1, 2), None), (None, (3, 38), None), (None, (39, 40), None)],
                    [(None, None, (1, 7)), (None, None, (8, 30))],
                    [(None, (1, 13), (1, 11)), (None, (14, 40), (1, 11)), (None, None, (12, 30))]):
        tot = np.zeros(len(ops))
        for w in windows:
            grp = ocb.FusedDiagnostics(*[ocb.Integral(o, indices=w) for o in ops]).compute()
            assert grp.plan.variant == 1
            tot += np.array(grp.integrals())
        for o, a, b in zip(ops, tot, ref):
            scale = float(ocb.Field(o).compute().interior.abs().sum()) / (64 * 40 * 30)
            assert abs(a - b) <= 1e-12 * scale, o.name


@pytest.mark.parametrize("size,halo,seed", [((64, 40, 30), (3, 3, 3), 1), ((70, 21, 9), (3, 3, 3), 2), ((96, 33, 70), (3, 3, 3), 3),
                                            ((64, 24, 12), (4, 4, 4), 4), ((38, 18, 40), (2, 2, 2), 5)], ids=str)
def test_class_sum_kernel_on_random_windows(size, halo, seed):
    """The class-sum kernel (integrals as weighted sums over faces / corners / edges, advection summed by parts with explicit
    window-edge terms) against the cell-exact integral mode of the budget kernel (OCB_FUSED_CELL_INTEGRALS=1) and the oracle's
    per-cell values summed over the window, for random y / z windows including single rows and planes, windows touching
    either end, partly filled x tiles (Nx = 70, 38), several z chunks (Nz = 70 with 32-plane chunks) and the even-halo tile
    pitch (36 columns, halo 2 and 4) next to the odd-halo one (34 columns)."""
    st = State(size=size, device="cuda", halo=halo)
    ops, oc = budget(st)
    want = oracle_budget(st, oc)
    Nx, Ny, Nz = size
    rng = np.random.default_rng(seed)
    windows = [(1, Ny, 1, Nz), (1, 1, 1, Nz), (Ny, Ny, Nz, Nz), (1, Ny, 1, 1), (2, Ny - 1, 2, Nz - 1)]
    for _ in range(4):
        j = sorted(rng.integers(1, Ny + 1, 2)); k = sorted(rng.integers(1, Nz + 1, 2))
        windows.append((int(j[0]), int(j[1]), int(k[0]), int(k[1])))
    os.environ["OCB_FUSED_KCHUNK"] = "32"
    try:
        for jlo, jhi, klo, khi in windows:
            idx = (None, (jlo, jhi), (klo, khi))
            sums = ocb.FusedDiagnostics(*[ocb.Integral(o, indices=idx) for o in ops]).compute()
            os.environ["OCB_FUSED_CELL_INTEGRALS"] = "1"
            try:
                cells = ocb.FusedDiagnostics(*[ocb.Integral(o, indices=idx) for o in ops]).compute()
            finally:
                os.environ.pop("OCB_FUSED_CELL_INTEGRALS", None)
            assert sums.plan.variant == 1 and cells.plan.variant == 1
            V = 1.0 / (Nx * Ny * Nz)
            for o, w_, a, b in zip(ops, want, sums.integrals(), cells.integrals()):
                win = w_[:, jlo - 1:jhi, klo - 1:khi]
                ref, scale = float(np.sum(win)) * V, float(np.sum(np.abs(win))) * V
                assert abs(a - ref) <= 1e-12 * scale, (o.name, (jlo, jhi, klo, khi), a, ref, scale)
                assert abs(a - b) <= 1e-12 * scale, (o.name, (jlo, jhi, klo, khi), a, b, scale)
    finally:
        os.environ.pop("OCB_FUSED_KCHUNK", None)


@pytest.mark.parametrize("size,halo,seed", [((64, 40, 30), (3, 3, 3), 11), ((70, 21, 9), (3, 3, 3), 12), ((96, 33, 70), (3, 3, 3), 13),
                                            ((64, 24, 12), (4, 4, 4), 14)], ids=str)
def test_class_sum_kernel_on_a_stretched_axis(size, halo, seed):
    """The class-sum kernel's stretched-z form (per-plane weights, the by-parts advection with the dz_c-weighted transports of
    the w equation) against the oracle's per-cell values times their cell volumes and against the cell-exact stretched kernel,
    for the whole domain and random y / z windows (single rows and planes, windows at either end, partly filled x tiles,
    several z chunks, both tile pitches)."""
    st = State(size=size, stretched=True, device="cuda", halo=halo)
    ops, oc = budget(st)
    want = oracle_budget(st, oc)
    Nx, Ny, Nz = size
    ii, jj, kk = st.og.indices("ccc")
    vol = np.broadcast_to(st.og.V("ccc", ii, jj, kk), want[0].shape)
    rng = np.random.default_rng(seed)
    windows = [(1, Ny, 1, Nz), (1, 1, 1, Nz), (Ny, Ny, Nz, Nz), (1, Ny, 1, 1), (2, Ny - 1, 2, Nz - 1)]
    for _ in range(4):
        j = sorted(rng.integers(1, Ny + 1, 2)); k = sorted(rng.integers(1, Nz + 1, 2))
        windows.append((int(j[0]), int(j[1]), int(k[0]), int(k[1])))
    os.environ["OCB_FUSED_KCHUNK"] = "32"
    try:
        for jlo, jhi, klo, khi in windows:
            idx = (None, (jlo, jhi), (klo, khi))
            sums = ocb.FusedDiagnostics(*[ocb.Integral(o, indices=idx) for o in ops]).compute()
            os.environ["OCB_DISABLE_SUMS_STRETCHED"] = "1"
            try:
                cells = ocb.FusedDiagnostics(*[ocb.Integral(o, indices=idx) for o in ops]).compute()
            finally:
                os.environ.pop("OCB_DISABLE_SUMS_STRETCHED", None)
            assert sums.plan.variant == 1 and cells.plan.variant == 1
            for o, w_, a, b in zip(ops, want, sums.integrals(), cells.integrals()):
                win, vw = w_[:, jlo - 1:jhi, klo - 1:khi], vol[:, jlo - 1:jhi, klo - 1:khi]
                ref, scale = float(np.sum(win * vw)), float(np.sum(np.abs(win) * vw))
                assert abs(a - ref) <= 1e-12 * scale, (o.name, (jlo, jhi, klo, khi), a, ref, scale)
                assert abs(a - b) <= 1e-12 * scale, (o.name, (jlo, jhi, klo, khi), a, b, scale)
    finally:
        os.environ.pop("OCB_FUSED_KCHUNK", None)


@pytest.mark.parametrize("stretched", [False, True], ids=["regular", "stretched"])
@pytest.mark.parametrize("size", [(64, 44, 26), (62, 23, 3), (34, 50, 9)], ids=str)
def test_ertel_pv_on_the_pipelined_kernel(size, stretched):
    """ErtelPotentialVorticity (fff, Nz + 1 faces) alone on the pipelined kernel: field, integral, and both (stretched z: the
    z differences over dz_f of their face, the fff point's volume dx dy dz_f)."""
    from cases import build_cases, check_case
    st = State(size=size, stretched=stretched, device="cuda")
    cases = {c[0]: c for c in build_cases(st, "scalar")}
    _, op, want, loc = cases["ErtelPotentialVorticity"]
    assert loc == "fff"
    for members in ([ocb.Field(op)], [ocb.Integral(op)], [ocb.Field(op), ocb.Integral(op)]):
        grp = ocb.FusedDiagnostics(*members).compute()
        torch.cuda.synchronize()
        assert grp.plan.variant == 1
        for m in grp.members:
            if isinstance(m, ocb.ComputedField):
                check_case(st, "ErtelPotentialVorticity", m.interior_ijk(), want)
        for integ in grp.integrals():
            assert abs(integ - K.integral(want, st.og, loc)) <= 1e-12 * integral_scale(st.og, want, loc)
    os.environ["OCB_DISABLE_FUSED"] = "1"
    try:
        gen = ocb.FusedDiagnostics(ocb.Field(op), ocb.Integral(op))
    finally:
        os.environ.pop("OCB_DISABLE_FUSED", None)
    gen.compute()
    assert gen.plan.variant == 0
    fused = ocb.FusedDiagnostics(ocb.Field(op), ocb.Integral(op)).compute()
    a, b = fused.members[0].interior_ijk(), gen.members[0].interior_ijk()
    assert a.shape == b.shape == (size[0], size[1], size[2] + 1)
    assert_close(a, b, 1e-11, "EPV fused vs generic")


def test_tke_budget_fields_on_a_stretched_axis():
    """The TKE budget about horizontal-mean profiles {SP, eps(u - U), PR, K} as FIELDS and integrals on a stretched z axis (the
    boundary-layer LES case): one launch of the budget kernel's per-plane-metric variant -- the mean shear of a cell is the mean
    of its two face gradients -- against the oracle."""
    from cases import build_cases, check_case
    st = State(size=(64, 44, 26), stretched=True, device="cuda")
    cases = {c[0]: c for c in build_cases(st, "scalar", mean="profile")}
    names = ["ShearProductionRate", "DissipationRate_perturbation", "KineticEnergyPressureRedistribution", "KineticEnergy"]
    ops = [cases[nm][1] for nm in names]
    grp = ocb.FusedDiagnostics(*([ocb.Field(o) for o in ops] + [ocb.Integral(o) for o in ops])).compute()
    torch.cuda.synchronize()
    assert grp.plan.variant == 1 and grp.plan.n_launches == 2
    vals = [m.interior_ijk() for m in grp.members if isinstance(m, ocb.ComputedField)]
    for nm, got, integ in zip(names, vals, grp.integrals()):
        _, _, want, loc = cases[nm]
        check_case(st, nm, got, want)
        assert abs(integ - K.integral(want, st.og, loc)) <= 1e-12 * integral_scale(st.og, want, loc), nm


@pytest.mark.parametrize("stretched", [False, True], ids=["regular", "stretched"])
def test_tke_budget_about_mean_profiles_one_sweep(stretched):
    """BASELINE.json configs[1] at test size: EPV + {ShearProductionRate, DissipationRate(u - U), PressureRedistribution}
    with U, V, W = horizontal-mean profiles, all in ONE launch of the pipelined kernel (perturbations about z profiles
    evaluated on the fly, EPV at the cell corners); values and integrals match the oracle.  Also on a stretched z axis."""
    from cases import build_cases, check_case
    st = State(size=(64, 44, 26), stretched=stretched, device="cuda")
    cases = {c[0]: c for c in build_cases(st, "scalar", mean="profile")}
    names = ["ShearProductionRate", "DissipationRate_perturbation", "KineticEnergyPressureRedistribution", "ErtelPotentialVorticity"]
    ops = [cases[nm][1] for nm in names]
    grp = ocb.FusedDiagnostics(*([ocb.Field(o) for o in ops] + [ocb.Integral(o) for o in ops])).compute()
    torch.cuda.synchronize()
    assert grp.plan.variant == 1 and grp.plan.n_launches == 2      # the sweep + the deterministic finalize of the partials
    vals = [m.interior_ijk() for m in grp.members if isinstance(m, ocb.ComputedField)]
    ints = grp.integrals()
    for nm, got, integ in zip(names, vals, ints):
        _, _, want, loc = cases[nm]
        check_case(st, nm, got, want)
        assert abs(integ - K.integral(want, st.og, loc)) <= 1e-12 * integral_scale(st.og, want, loc), nm
    # integrals only, and the z shear production alone equals the total for profile means
    grp2 = ocb.FusedDiagnostics(*[ocb.Integral(cases[nm][1]) for nm in names[:3]] + [ocb.Integral(cases["ZShearProductionRate"][1])])
    # SP and SPZ in one plan: one SP request per pipelined sweep, so this plan runs on the generic kernel
    grp2.compute()
    i2 = grp2.integrals()
    for a, b, nm in zip(i2[:3], ints[:3], names[:3]):
        assert abs(a - b) <= 1e-12 * integral_scale(st.og, cases[nm][2], "ccc"), nm
    assert abs(i2[3] - i2[0]) <= 1e-12 * integral_scale(st.og, cases["ShearProductionRate"][2], "ccc")
    only = ocb.FusedDiagnostics(*[ocb.Integral(cases[nm][1]) for nm in names[:3]])
    assert only.plan.variant == 1
    for a, b, nm in zip(only.compute().integrals(), ints[:3], names[:3]):
        assert abs(a - b) <= 1e-12 * integral_scale(st.og, cases[nm][2], "ccc"), nm


@pytest.mark.parametrize("stretched", [False, True], ids=["regular", "stretched"])
@pytest.mark.parametrize("size,seed", [((64, 44, 26), 1), ((70, 21, 40), 2), ((96, 48, 70), 3)], ids=str)
def test_class_sum_kernel_tke_ke_and_tracer_variance_budget(size, seed, stretched):
    """north_star's target plan -- the TKE + KE + tracer-variance budget as volume integrals -- in ONE launch of the class-sum
    kernel: the seven KE / tracer-variance terms plus, about the horizontal-mean profiles U(z), V(z), W(z), the shear production
    (summed over x / y / z faces), the dissipation of u - U and the turbulent kinetic energy.  Against the oracle's per-cell
    values summed over the window, and against the generic kernel, on the whole domain and on random y / z windows."""
    from cases import build_cases
    st = State(size=size, stretched=stretched, device="cuda")      # (stretched z: the per-plane-weight form of the kernel)
    cases = {c[0]: c for c in build_cases(st, "scalar", mean="profile")}
    names = ["KineticEnergy", "KineticEnergyAdvection", "KineticEnergyPressureRedistribution", "PotentialEnergyConversion", "DissipationRate",
             "TracerVarianceDissipationRate", "TracerVarianceDiffusion", "ShearProductionRate", "DissipationRate_perturbation",
             "TurbulentKineticEnergy"]
    ops = [cases[nm][1] for nm in names]
    Nx, Ny, Nz = size
    ii, jj, kk = st.og.indices("ccc")
    V = np.broadcast_to(st.og.V("ccc", ii, jj, kk), cases[names[0]][2].shape)      # cell volumes
    rng = np.random.default_rng(seed)
    windows = [None, (1, Ny, 1, 1), (1, 1, 1, Nz), (2, Ny - 1, 2, Nz - 1), (Ny, Ny, Nz, Nz)]
    for _ in range(3):
        j = sorted(rng.integers(1, Ny + 1, 2)); k = sorted(rng.integers(1, Nz + 1, 2))
        windows.append((int(j[0]), int(j[1]), int(k[0]), int(k[1])))
    os.environ["OCB_FUSED_KCHUNK"] = "32"
    try:
        for win in windows:
            idx = None if win is None else (None, (win[0], win[1]), (win[2], win[3]))
            jlo, jhi, klo, khi = win or (1, Ny, 1, Nz)
            sums = ocb.FusedDiagnostics(*[ocb.Integral(o, indices=idx) for o in ops]).compute()
            assert sums.plan.variant == 1 and sums.plan.n_launches == 2, (win, sums.plan.variant)
            os.environ["OCB_DISABLE_FUSED"] = "1"
            try:
                gen = ocb.FusedDiagnostics(*[ocb.Integral(o, indices=idx) for o in ops]).compute()
            finally:
                os.environ.pop("OCB_DISABLE_FUSED", None)
            assert gen.plan.variant == 0
            for nm, a, b in zip(names, sums.integrals(), gen.integrals()):
                w_, v_ = cases[nm][2][:, jlo - 1:jhi, klo - 1:khi], V[:, jlo - 1:jhi, klo - 1:khi]
                ref, scale = float(np.sum(w_ * v_)), float(np.sum(np.abs(w_) * v_))
                assert abs(a - ref) <= 1e-12 * scale, (nm, win, a, ref, scale)
                assert abs(a - b) <= 1e-12 * scale, (nm, win, a, b, scale)
    finally:
        os.environ.pop("OCB_FUSED_KCHUNK", None)
    # the TKE-budget subset alone (config 2 without the potential vorticity) has its own compiled variant
    sub = ["ShearProductionRate", "DissipationRate_perturbation", "KineticEnergyPressureRedistribution", "TurbulentKineticEnergy"]
    g = ocb.FusedDiagnostics(*[ocb.Integral(cases[nm][1]) for nm in sub]).compute()
    assert g.plan.variant == 1
    for nm, a in zip(sub, g.integrals()):
        w_ = cases[nm][2]
        assert abs(a - float(np.sum(w_ * V))) <= 1e-12 * float(np.sum(np.abs(w_) * V)), nm
    # no mean flow at all (ZeroField means): TKE is K
    cz = {c[0]: c for c in build_cases(st, "scalar", mean="none")}
    gz = ocb.FusedDiagnostics(*[ocb.Integral(cz[nm][1]) for nm in ("TurbulentKineticEnergy", "KineticEnergy")]).compute()
    a, b = gz.integrals()
    assert abs(a - b) <= 1e-13 * abs(b)


def test_split_plan_budget_kernel_plus_generic_kernel():
    """Requests outside the pipelined kernel's set (Richardson number), or outside a variant's envelope (EPV next to the
    advection term), go to the generic kernel of the same plan (variant 2) and fill their own integral slots."""
    from cases import build_cases
    st = State(size=(64, 44, 26), device="cuda")
    ops, oc = budget(st)
    cases = {c[0]: c for c in build_cases(st, "scalar")}
    extra = ["ErtelPotentialVorticity", "RossbyNumber"]
    grp = ocb.FusedDiagnostics(*[ocb.Integral(o) for o in ops + [cases[nm][1] for nm in extra]]).compute()
    assert grp.plan.variant == 2 and grp.plan.n_launches == 3
    ints = grp.integrals()
    for o, w_, integ in zip(ops, oracle_budget(st, oc), ints[:len(ops)]):
        assert abs(integ - K.integral(w_, st.og, "ccc")) <= 1e-12 * integral_scale(st.og, w_, "ccc"), o.name
    for nm, integ in zip(extra, ints[len(ops):]):
        _, _, want, loc = cases[nm]
        assert abs(integ - K.integral(want, st.og, loc)) <= 1e-12 * integral_scale(st.og, want, loc), nm


def test_full_size_1024_cube_properties():
    """BASELINE.json configs[2] at its full size (1024^3 Float64, the bench workload), through properties that need no
    oracle: run-to-run bit equality; the pipelined kernel against the per-cell generic kernel (two independent CUDA
    formulations) to 1e-12 of each integrand's scale; windows of the domain against the oracle's C twin; exact power-of-two scaling (u -> 2u multiplies K and eps by 4, ADV by
    8, PR and u_i b_i by 2, bit for bit); y windows add up to the whole; <chi> = <2c div F>."""
    import sys
    from helpers import ROOT
    free, _ = torch.cuda.mem_get_info()
    if free < 70e9:
        pytest.skip("needs ~60 GB of device memory")
    sys.path.insert(0, ROOT)
    import bench
    n = 1024
    grid, m = bench.make_state(ocb, torch, (n, n, n), torch.device("cuda", 0))
    ops = bench.budget_ops(ocb, m)
    fused = ocb.FusedDiagnostics(*[ocb.Integral(o) for o in ops])
    assert fused.plan.variant == 1
    first = fused.compute().integrals()
    assert fused.compute().integrals() == first
    # generic kernel: one request at a time keeps its scratch small; scale = the integral of |integrand|
    os.environ["OCB_DISABLE_FUSED"] = "1"
    try:
        generic = [ocb.FusedDiagnostics(ocb.Integral(o)) for o in ops]
    finally:
        os.environ.pop("OCB_DISABLE_FUSED", None)
    dV = 1.0 / n ** 3
    for o, a, g in zip(ops, first, generic):
        assert g.plan.variant == 0
        b = g.compute().integrals()[0]
        f = ocb.Field(o).compute()
        scale = float(f.interior.abs().sum()) * dV
        del f
        torch.cuda.empty_cache()
        assert abs(a - b) <= 1e-12 * scale, (o.name, a, b, scale)
    assert abs(first[5] - first[6]) <= 1e-10 * abs(first[5])                      # <chi> = <2c div F>
    halves = [ocb.FusedDiagnostics(*[ocb.Integral(o, indices=(None, w, None)) for o in ops]).compute().integrals()
              for w in ((1, 511), (512, n))]
    for o, a, h0, h1 in zip(ops, first, *halves):
        assert abs((h0 + h1) - a) <= 1e-11 * max(abs(h0), abs(h1), abs(a)), o.name
    # against the ORACLE at full size: the class-sum integrals over a 1024 x 64 x 64 window (whole x rows, as the kernel
    # requires), against the oracle's C twin evaluated per cell on the same block of the parent arrays (its halo cells are
    # the window's real neighbours), to 1e-12 of each integrand's scale
    import __graft_entry__ as ge
    from oracle import cbudget
    ge.build_oracle()
    H, nw = 3, 64
    for j0, k0 in ((481, 1), (1, 961), (700, 333)):
        win = ocb.FusedDiagnostics(*[ocb.Integral(o, indices=(None, (j0, j0 + nw - 1), (k0, k0 + nw - 1))) for o in ops])
        assert win.plan.variant == 1
        got = win.compute().integrals()
        blk = lambda f, nz: f.data[k0 - 1:k0 - 1 + nz + 2 * H, j0 - 1:j0 - 1 + nw + 2 * H, :].cpu().numpy()
        stt = cbudget.BudgetState((n, nw, nw), H, (1.0 / n,) * 3, 1e-6, 1e-7, blk(m.velocities.u, nw), blk(m.velocities.v, nw),
                                  blk(m.velocities.w, nw + 1), blk(m.tracers.b, nw), blk(m.pressures.pNHS, nw))
        for o, d, a in zip(ops, cbudget.DIAGS, got):
            ref = stt.integral(d)
            scale = float(np.abs(stt.field(d)).sum()) * dV
            assert abs(a - ref) <= 1e-12 * scale, (o.name, (j0, k0), a, ref, scale)
    for f in (m.velocities.u, m.velocities.v, m.velocities.w):
        f.data.mul_(2.0)
    scaled = fused.compute().integrals()
    for a, b, factor in zip(scaled, first, (4.0, 8.0, 2.0, 2.0, 4.0, 1.0, 1.0)):
        assert a == factor * b
