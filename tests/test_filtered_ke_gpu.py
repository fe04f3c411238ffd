"""GPU parity tests of the filtered / sub-filter kinetic-energy diagnostics (BASELINE.json configs[4] at test size):
the staged Gaussian filter (K2) feeding the contraction kernels (K1), against the same graph assembled from the NumPy
oracle (reference graph: src/FilteredKineticEnergyEquation.jl:29-48, 176-188, 275-299, 366-389;
src/SubFilterKineticEnergyEquation.jl:29, 76-85, 143-151; cross-checks test/test_filtered_kinetic_energy_equation.jl:52-84,
test/test_subfilter_kinetic_energy_equation.jl:16-41)."""
import numpy as np
import pytest
import torch

from helpers import State, K, ocb, assert_close
from oracle.grid import materialize
from oracle import filters as OF

pytestmark = pytest.mark.gpu
FKE, SFKE, SF = ocb.FilteredKineticEnergyEquation, ocb.SubFilterKineticEnergyEquation, ocb.SpatialFilters

SIGMA, DIMS, BOUNDARY = 0.09, (1, 2, 3), "shrink"


class Lazy:
    def __init__(self, fn):
        self.fn = fn

    def __getitem__(self, ijk):
        return self.fn(*ijk)


def ofilt(st, values, loc):
    """materialised Field(filter(Field(values))) on the oracle side."""
    f = materialize(st.og, loc, values)
    return materialize(st.og, loc, OF.gaussian_filter(f, DIMS, SIGMA, None, BOUNDARY))


def oracle_graph(st, oc):
    og = st.og
    ub, vb, wb = ofilt(st, st.ou.interior, "fcc"), ofilt(st, st.ov.interior, "cfc"), ofilt(st, st.ow.interior, "ccf")
    flds = st.ofields()
    fl = lambda f, loc: ofilt(st, K.evaluate(f, og, loc, oc, None, None, flds, None), loc)
    ff = {"t11": fl(K.viscous_flux_ux, "ccc"), "t22": fl(K.viscous_flux_vy, "ccc"), "t33": fl(K.viscous_flux_wz, "ccc"),
          "t12": fl(K.viscous_flux_uy, "ffc"), "t13": fl(K.viscous_flux_uz, "fcf"), "t23": fl(K.viscous_flux_vz, "cff")}
    ff.update(t21=ff["t12"], t31=ff["t13"], t32=ff["t23"])
    fv = {"u": ub, "v": vb, "w": wb}
    ke = K.evaluate(K.kinetic_energy_ccc, og, "ccc", st.ou, st.ov, st.ow)
    eps = K.evaluate(K.viscous_dissipation_rate_ccc, og, "ccc", None, flds, {"closure": oc})
    out = {"Kl": K.evaluate(K.kinetic_energy_ccc, og, "ccc", ub, vb, wb),
           "epsl": K.evaluate(K.filtered_dissipation_rate_ccc, og, "ccc", fv, ff)}
    out["Ks"] = ofilt(st, ke, "ccc").interior - out["Kl"]
    out["epss"] = ofilt(st, eps, "ccc").interior - out["epsl"]
    # cross-scale flux: -Σ w_ij <τ_ij>_c <S̄_ij>_c, τ_ij = filter(u_i u_j) - ū_i ū_j at the natural locations
    comp = {(1, 1): ("fcc", (st.ou,), (ub,)), (2, 2): ("cfc", (st.ov,), (vb,)), (3, 3): ("ccf", (st.ow,), (wb,)),
            (1, 2): ("ffc", (st.ou, st.ov), (ub, vb)), (1, 3): ("fcf", (st.ou, st.ow), (ub, wb)), (2, 3): ("cff", (st.ov, st.ow), (vb, wb))}
    sloc = {(1, 1): "ccc", (2, 2): "ccc", (3, 3): "ccc", (1, 2): "ffc", (1, 3): "fcf", (2, 3): "cff"}
    i, j, k = og.indices("ccc")
    pi = 0.0
    for (a, b), (loc, full_args, filt_args) in comp.items():
        full = ofilt(st, K.evaluate(K.stress_tensor_component(a, b), og, loc, *full_args), loc)
        sfn = K.stress_tensor_component(a, b)
        tau = Lazy(lambda ii, jj, kk, full=full, sfn=sfn, fa=filt_args: full[ii, jj, kk] - sfn(ii, jj, kk, og, *fa))
        strain = K.strain_rate_tensor_component(a, b)
        S = Lazy(lambda ii, jj, kk, strain=strain, fa=filt_args: strain(ii, jj, kk, og, *fa))
        wgt = 1 if a == b else 2
        pi = pi + wgt * K.to_center(i, j, k, og, tau, loc) * K.to_center(i, j, k, og, S, sloc[(a, b)])
    out["Pi"] = np.broadcast_to(-pi, og.extent_of("ccc")).copy()
    return out


def test_filtered_and_subfilter_ke_terms_match_oracle_graph():
    st = State(size=(20, 16, 12), device="cuda")
    oc, pc = st.closures()["scalar"]
    m = st.model(closure=pc)
    filt = SF.GaussianFilter(dims=DIMS, sigma=SIGMA, boundary=BOUNDARY)
    ops = {"Kl": FKE.FilteredKineticEnergy(m, filt), "epsl": FKE.FilteredKineticEnergyDissipationRate(m, filt),
           "Ks": SFKE.SubFilterKineticEnergy(m, filt), "epss": SFKE.SubFilterKineticEnergyDissipationRate(m, filt),
           "Pi": FKE.KineticEnergyCrossScaleFlux(m, filt)}
    want = oracle_graph(st, oc)
    for name, op in ops.items():
        got = ocb.Field(op).compute().interior_ijk()
        torch.cuda.synchronize()
        assert_close(got, want[name], 1e-11, name)


def test_decomposition_identity_and_recompute():
    """filter(K) = Kˡ + Kˢ holds by construction (SubFilterKineticEnergyEquation.jl:22-27); recomputation follows the
    evolving model fields (test/test_subfilter_kinetic_energy_equation.jl:61-75)."""
    st = State(size=(16, 16, 8), device="cuda")
    m = st.model(closure=st.closures()["scalar"][1])
    filt = SF.GaussianFilter(dims=(1, 2), sigma=0.1)
    kl, ks = ocb.Field(FKE.FilteredKineticEnergy(m, filt)), ocb.Field(SFKE.SubFilterKineticEnergy(m, filt))
    kbar = ocb.Field(filt(ocb.Field(ocb.KineticEnergyEquation.KineticEnergy(m))))
    for t in (1.0, 2.0):
        for f in (kl, ks, kbar):
            f.compute(time=t)
        torch.cuda.synchronize()
        a, b = (kl.interior + ks.interior).cpu().numpy(), kbar.interior.cpu().numpy()
        np.testing.assert_allclose(a, b, rtol=1e-13, atol=1e-16)
        st.u.data.mul_(1.5)          # evolve the model state: the next compute must see it
    assert float(kl.interior.mean()) > 0


def test_cross_scale_flux_dims_subset_and_collocated_diagonals():
    st = State(size=(16, 12, 10), device="cuda")
    m = st.model(closure=st.closures()["scalar"][1])
    filt = SF.GaussianFilter(dims=(1, 2), sigma=0.1)
    full = ocb.Field(FKE.KineticEnergyCrossScaleFlux(m, filt)).compute().interior_ijk()
    xz = ocb.Field(FKE.KineticEnergyCrossScaleFlux(m, filt, dims=(1, 3))).compute().interior_ijk()
    col = ocb.Field(FKE.KineticEnergyCrossScaleFlux(m, filt, collocate_diagonals=True)).compute().interior_ijk()
    torch.cuda.synchronize()
    assert np.isfinite(full).all() and np.isfinite(xz).all() and np.isfinite(col).all()
    assert not np.allclose(full, xz) and not np.allclose(full, col)
    with pytest.raises(ValueError):
        FKE.KineticEnergyCrossScaleFlux(m, filt, dims=(1, 1))


def test_subfilter_covariance_and_stress_tensor_against_the_oracle_graph():
    """subfilter_covariance(a, b, filter) = filter(a b) - filter(a) filter(b) with both operands co-located first
    (src/FlowDiagnostics.jl:901-910) and subfilter_stress_tensor component by component (src/FilteredKineticEnergyEquation.jl:42-48),
    against the same graphs assembled from the NumPy oracle; the result recomputes its whole graph on compute()."""
    st = State(size=(24, 20, 16), device="cuda")
    og = st.og
    FD = ocb.FlowDiagnostics
    filt = SF.GaussianFilter(dims=DIMS, sigma=SIGMA, boundary=BOUNDARY)
    m = st.model(closure=st.closures()["scalar"][1])
    # tracer flux carried by the sub-filter scales: a = u (Face in x), b = the tracer (Center)
    cov = FD.subfilter_covariance(m.velocities.u, m.tracers.b, filt)
    cov.compute()
    torch.cuda.synchronize()
    uc = materialize(og, "ccc", K.evaluate(K.ix_caa, og, "ccc", st.ou))                 # @at ccc u
    fu, fb = ofilt(st, uc.interior, "ccc"), ofilt(st, st.ob.interior, "ccc")
    fub = ofilt(st, uc.interior * st.ob.interior, "ccc")
    want = fub.interior - fu.interior * fb.interior
    assert_close(cov.interior_ijk(), want, 1e-12, "subfilter_covariance(u, b)")
    # a shared pre-filtered first factor gives the same field
    shared = FD.subfilter_covariance(m.velocities.u, m.tracers.b, filt, filtered_a=ocb.Field(filt(FD.to_center(m.velocities.u)))).compute()
    np.testing.assert_array_equal(shared.interior_ijk(), cov.interior_ijk())
    # the graph is live: change the tracer, recompute, compare again
    m.tracers.b.data.mul_(-2.0)
    cov.compute(time=1.0)
    assert_close(cov.interior_ijk(), -2.0 * want, 1e-12, "subfilter_covariance after an update")
    m.tracers.b.data.mul_(-0.5)
    # sub-filter stress, off-diagonal component at its own (Face, Face, Center) location
    tau = FKE.subfilter_stress_tensor(m, filt)
    assert sorted(vars(tau)) == sorted(["τ₁₁", "τ₂₂", "τ₃₃", "τ₁₂", "τ₁₃", "τ₂₃"])
    t12 = getattr(tau, "τ₁₂").compute()
    ub, vb = ofilt(st, st.ou.interior, "fcc"), ofilt(st, st.ov.interior, "cfc")
    full = ofilt(st, K.evaluate(K.stress_tensor_component(1, 2), og, "ffc", st.ou, st.ov), "ffc")
    want12 = full.interior - K.evaluate(K.stress_tensor_component(1, 2), og, "ffc", ub, vb)
    assert t12.location == ("f", "f", "c")
    assert_close(t12.interior_ijk(), want12, 1e-12, "τ₁₂")


def test_names_outside_the_hot_path_fail_loudly():
    st = State(size=(8, 8, 8), device="cuda")
    m = st.model(closure=st.closures()["scalar"][1])
    for fn in (ocb.FlowDiagnostics.MixedLayerDepth,):
        with pytest.raises(ocb.OcbError, match="Oceanostics.jl"):
            fn(m)
    # the kinetic-energy tendency is evaluated for the plain NonhydrostaticModel; what the library does not restate is refused
    m.advection = "WENO(order=5)"
    with pytest.raises(ocb.OcbError, match="Centered"):           # (Centered(order = 4) IS evaluated: cases "_o4")
        ocb.KineticEnergyEquation.KineticEnergyTendency(m)
    m.advection, m.stokes_drift = "Centered", object()
    with pytest.raises(ocb.OcbError, match="stokes_drift"):
        ocb.KineticEnergyEquation.KineticEnergyTendency(m)
    assert ocb.KineticEnergyEquation.PotentialToKineticEnergyConversion is ocb.KineticEnergyEquation.PotentialEnergyConversion
    assert SFKE.KineticEnergyCrossScaleFlux is FKE.KineticEnergyCrossScaleFlux
