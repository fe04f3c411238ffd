"""GPU parity tests of the velocity-gradient kernel (csrc/ocb_sweep_strain.cu): closure-dependent DissipationRate,
isotropic dissipation, StrainRateTensorModulus, VorticityTensorModulus and Q in one sweep, Float64 (rtol 1e-12) and
Float32 (1e-5 relative L2), regular and stretched z, constant / eddy-viscosity-field / tuple closures -- against the
NumPy oracle and against the generic kernel."""
import os

import numpy as np
import pytest
import torch

from helpers import State, assert_close, K, ocb
from cases import build_cases, check_case, integral_scale

pytestmark = pytest.mark.gpu

NAMES = ["DissipationRate", "KineticEnergyIsotropicDissipationRate", "StrainRateTensorModulus", "VorticityTensorModulus",
         "QVelocityGradientTensorInvariant"]


def _group(ops, fields=True, integrals=True, generic=False):
    if generic:
        os.environ["OCB_DISABLE_FUSED"] = "1"
    try:
        grp = ocb.FusedDiagnostics(*(([ocb.Field(o) for o in ops] if fields else []) + ([ocb.Integral(o) for o in ops] if integrals else [])))
    finally:
        os.environ.pop("OCB_DISABLE_FUSED", None)
    grp.compute()
    torch.cuda.synchronize()
    return grp


@pytest.mark.parametrize("stretched", [False, True], ids=["regular", "stretched"])
@pytest.mark.parametrize("closure", ["scalar", "smagorinsky_like", "tuple"])
@pytest.mark.parametrize("size", [(40, 24, 18), (33, 17, 5), (64, 50, 37)], ids=str)
def test_velocity_gradient_family_matches_oracle_f64(size, closure, stretched):
    st = State(size=size, stretched=stretched, device="cuda")
    cases = {c[0]: c for c in build_cases(st, closure)}
    ops = [cases[nm][1] for nm in NAMES]
    grp = _group(ops)
    # rows come in as aligned element pairs: an odd Nx (odd parent row length) takes the generic kernel
    kv = 3 if size[0] % 2 == 0 else 0
    assert grp.plan.variant == kv and grp.plan.n_launches == 2
    vals = [m.interior_ijk() for m in grp.members if isinstance(m, ocb.ComputedField)]
    for nm, got, integ in zip(NAMES, vals, grp.integrals()):
        _, _, want, loc = cases[nm]
        check_case(st, nm, got, want)
        assert abs(integ - K.integral(want, st.og, loc)) <= 1e-12 * integral_scale(st.og, want, loc), nm
    # subsets compile to their own edge families: each alone, integrals only
    for nm in NAMES:
        one = _group([cases[nm][1]], fields=False)
        budget_takes_it = nm == "DissipationRate" and closure == "scalar" and size[0] % 2 == 0     # regular or stretched z
        assert one.plan.variant == (1 if budget_takes_it else kv)
        want, loc = cases[nm][2], cases[nm][3]
        assert abs(one.integrals()[0] - K.integral(want, st.og, loc)) <= 1e-12 * integral_scale(st.og, want, loc), nm


def test_float32_config4_shape_and_generic_agreement():
    """BASELINE.json configs[3] at test size: Float32, stretched z, eddy-viscosity closures: eps + |S| on the
    velocity-gradient kernel, to 1e-5 relative L2 of the Float64 oracle and of the generic kernel."""
    st = State(size=(96, 70, 40), stretched=True, dtype=np.float32, device="cuda")
    for closure in ("smagorinsky_like", "tuple", "scalar"):
        oc, pc = st.closures()[closure]
        m = st.model(closure=pc)
        ops = [ocb.KineticEnergyEquation.DissipationRate(m), ocb.FlowDiagnostics.StrainRateTensorModulus(m)]
        grp = _group(ops)
        assert grp.plan.variant == 3
        gen = _group(ops, generic=True)
        assert gen.plan.variant == 0
        want = [K.evaluate(K.viscous_dissipation_rate_ccc, st.og, "ccc", None, st.ofields(), {"closure": oc}),
                K.evaluate(K.strain_rate_tensor_modulus_ccc, st.og, "ccc", st.ou, st.ov, st.ow)]
        for a, b, w_, ia, ib in zip(grp.members[:2], gen.members[:2], want, grp.integrals(), gen.integrals()):
            ga, gb, w64 = a.interior_ijk().astype(np.float64), b.interior_ijk().astype(np.float64), w_.astype(np.float64)
            assert np.linalg.norm(ga - w64) <= 1e-5 * np.linalg.norm(w64)
            assert np.linalg.norm(ga - gb) <= 1e-5 * np.linalg.norm(gb)
            assert abs(ia - K.integral(w64, st.og, "ccc")) <= 1e-5 * integral_scale(st.og, w64, "ccc")
            assert abs(ia - ib) <= 1e-5 * abs(ib)


def test_windows_and_split_plans():
    """y / z windows add up to the whole-domain integrals; a plan that also holds other diagnostics is split between the
    budget kernel, the velocity-gradient kernel and the generic kernel."""
    st = State(size=(64, 40, 30), stretched=True, device="cuda")
    cases = {c[0]: c for c in build_cases(st, "smagorinsky_like")}
    ops = [cases[nm][1] for nm in NAMES]
    ref = _group(ops, fields=False).integrals()
    for windows in ([(None, (1, 2), None), (None, (3, 38), None), (None, (39, 40), None)],
                    [(None, None, (1, 7)), (None, None, (8, 30))]):
        tot = np.zeros(len(ops))
        for w in windows:
            grp = ocb.FusedDiagnostics(*[ocb.Integral(o, indices=w) for o in ops]).compute()
            assert grp.plan.variant == 3
            tot += np.array(grp.integrals())
        for nm, a, b in zip(NAMES, tot, ref):
            assert abs(a - b) <= 1e-12 * integral_scale(st.og, cases[nm][2], "ccc"), nm
    # stretched z with an eddy-viscosity closure: |S| and Q on the velocity-gradient kernel, the closure-independent KE on the
    # budget kernel's per-plane-metric variant, Ro on the generic one
    mixed = ["StrainRateTensorModulus", "KineticEnergy", "QVelocityGradientTensorInvariant", "RossbyNumber"]
    grp = ocb.FusedDiagnostics(*[ocb.Integral(cases[nm][1]) for nm in mixed]).compute()
    assert grp.plan.variant == 6 and grp.plan.n_launches == 4
    for nm, integ in zip(mixed, grp.integrals()):
        _, _, want, loc = cases[nm]
        assert abs(integ - K.integral(want, st.og, loc)) <= 1e-12 * integral_scale(st.og, want, loc), nm
    # regular z, constant closure: budget kernel (K, eps) + velocity-gradient kernel (|S|) + generic (Ro) in one plan
    st2 = State(size=(64, 40, 30), device="cuda")
    c2 = {c[0]: c for c in build_cases(st2, "scalar")}
    names = ["KineticEnergy", "DissipationRate", "StrainRateTensorModulus", "RossbyNumber"]
    grp = ocb.FusedDiagnostics(*[ocb.Integral(c2[nm][1]) for nm in names]).compute()
    assert grp.plan.variant == 6 and grp.plan.n_launches == 4
    for nm, integ in zip(names, grp.integrals()):
        _, _, want, loc = c2[nm]
        assert abs(integ - K.integral(want, st2.og, loc)) <= 1e-12 * integral_scale(st2.og, want, loc), nm


@pytest.mark.parametrize("stretched", [False, True], ids=["regular", "stretched"])
def test_smagorinsky_viscosity_computed_on_the_device_feeds_the_dissipation(stretched):
    """with closure = SmagorinskyLilly the eddy viscosity is a computed Field of the library
    (OCB_DIAG_NU_SMAG) that every diagnostic refreshes first -- no nu_e array is handed in.  eps and |S| then run on the
    velocity-gradient kernel and must equal the oracle's eps with the oracle's own Smagorinsky viscosity."""
    from oracle.grid import OField
    st = State(size=(40, 24, 18), stretched=stretched, device="cuda")
    sm = ocb.SmagorinskyLilly(C=0.16, Cb=1.0, Pr=0.8)
    m = st.model(closure=sm)
    assert m.closure_fields.nu_e is sm.nu_e and sm.kappa_scale == (1.25,)
    eps = ocb.Field(ocb.KineticEnergyEquation.DissipationRate(m))
    smod = ocb.Field(ocb.FlowDiagnostics.StrainRateTensorModulus(m))
    grp = ocb.FusedDiagnostics(eps, smod)
    grp.compute(time=1.0)
    torch.cuda.synchronize()
    assert grp.plan.variant == 3
    og = st.og
    onu = OField(og, "ccc")
    onu.interior[...] = K.evaluate(K.smagorinsky_nu_ccc, og, "ccc", st.ou, st.ov, st.ow, st.ob, 0.16, 1.0, 0.8)
    onu.fill_halo()
    np.testing.assert_allclose(sm.nu_e.interior_ijk(), onu.interior, rtol=1e-12, atol=1e-22)
    want = K.evaluate(K.viscous_dissipation_rate_ccc, og, "ccc", None, st.ofields(), {"closure": K.Closure(nu=onu, Pr=0.8)})
    check_case(st, "DissipationRate", eps.interior_ijk(), want)
    # the state evolves: the next compute refreshes the viscosity before the sweep
    st.u.data.mul_(1.5)
    grp.compute(time=2.0)
    torch.cuda.synchronize()
    assert float(sm.nu_e.interior.max()) > 0 and not np.allclose(sm.nu_e.interior_ijk(), onu.interior)
    assert m.compute_diffusivities(time=3.0).nu_e is sm.nu_e
