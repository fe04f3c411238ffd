"""GPU tests of the available-potential-energy diagnostics (scope table 8f, row 1: the consumers of the K3 sort) through the
package's constructors -- sort, Ψδ and the pointwise / flux kernels end to end -- against the reference's own analytic
cases (test/test_ape_diagnostics.jl:85-122, 541-563, 565-608) and the NumPy oracle."""
import numpy as np
import pytest
import torch

from helpers import ocb, K
from oracle import bpe as OB
from oracle.grid import OField
from test_bpe_gpu import _pair, ztol

pytestmark = pytest.mark.gpu
APE, BPE, TV = ocb.AvailablePotentialEnergyEquation, ocb.BackgroundPotentialEnergyEquation, ocb.TracerVarianceEquation


def _model(pg, pb, closure=None):
    m = ocb.NonhydrostaticModel(pg, tracers=("b",), closure=closure)
    m.tracers.b = pb
    return m


def _integral(op):
    return ocb.Field(ocb.Integral(op)).compute().value()


def test_two_cell_column_known_answers():
    """:85-107: dense over light: ∫E_b = -0.25, ∫E_a = +0.50; flipped (stable): ∫E_a = 0."""
    og, pg, ob, pb = _pair((1, 1, 2), (-1, 0), np.array([1.0, -1.0]).reshape(1, 1, 2))
    m = _model(pg, pb)
    z = APE.reference_height(m)
    np.testing.assert_allclose(z.interior_ijk().ravel(), [-0.25, -0.75], rtol=1e-14)
    assert _integral(APE.BackgroundPotentialEnergy(m, z)) == pytest.approx(-0.25, abs=1e-14)
    assert _integral(APE.AvailablePotentialEnergy(m, z)) == pytest.approx(0.50, abs=1e-14)
    pb.interior.copy_(torch.tensor([-1.0, 1.0], dtype=torch.float64, device="cuda").reshape(2, 1, 1))
    pb.fill_halo_regions()
    z = APE.reference_height(m)
    np.testing.assert_allclose(z.interior_ijk().ravel(), [-0.75, -0.25], rtol=1e-14)
    assert abs(_integral(APE.AvailablePotentialEnergy(m, z))) <= 1e-14


@pytest.mark.parametrize("method", ["ThreeDimensionalSort", "HeavisideIntegral"])
def test_sorted_state_holds_no_available_energy(method):
    """:108-122 and :592-608: b = 3z is its own sorted state: ∫E_a ~ 0, Υ ~ 0, ε_a ~ 0 against χ/2."""
    Nx, Ny, Nz = 16, 12, 20
    zc = -1 + (np.arange(Nz) + 0.5) / Nz
    og, pg, ob, pb = _pair((Nx, Ny, Nz), (-1, 0), np.broadcast_to(3 * zc, (Nx, Ny, Nz)))
    m = _model(pg, pb, closure=ocb.Closure(nu=1e-6, kappa=1e-3))
    meth = getattr(BPE, method)()
    Ep = float(np.sum(-ob.interior * zc.reshape(1, 1, Nz)) / (Nx * Ny * Nz))
    root = np.sqrt(np.finfo(np.float64).eps)
    assert abs(_integral(APE.AvailablePotentialEnergy(m, method=meth))) <= root * abs(Ep)
    assert _integral(APE.BackgroundPotentialEnergy(m, method=meth)) == pytest.approx(Ep, rel=1e-12)
    if method == "HeavisideIntegral":            # tied planes: only the tie-free method puts every cell at its own height
        ups = ocb.Field(APE.BuoyancyDisplacementPotential(m)).compute().interior_ijk()
        assert np.abs(ups).max() < root
        chi = ocb.Field(TV.TracerVarianceDissipationRate(m, "b")).compute().interior_ijk()
        eps_a = ocb.Field(APE.AvailablePotentialEnergyDissipationRate(m)).compute().interior_ijk()
        assert chi.max() > 0 and np.abs(eps_a).max() < root * chi.max() / 2


@pytest.mark.parametrize("stretched", [False, True], ids=["regular", "stretched"])
def test_random_field_against_oracle_and_identities(stretched):
    from helpers import stretched_faces
    rng = np.random.default_rng(17)
    Nx, Ny, Nz = 20, 18, 16
    vals = rng.standard_normal((Nx, Ny, Nz))
    zspec = stretched_faces(Nz) if stretched else (-1, 0)
    og, pg, ob, pb = _pair((Nx, Ny, Nz), zspec, vals)
    oc = K.Closure(nu=1e-6, kappa=1e-3)
    m = _model(pg, pb, closure=ocb.Closure(nu=1e-6, kappa=1e-3))
    z = APE.reference_height(m, method=BPE.HeavisideIntegral())
    want = OB.reference_height(ob, "HeavisideIntegral")

    def ofield(v):
        f = OField(og, "ccc")
        f.interior[...] = v
        f.fill_halo()
        return f
    ozs, opsi = ofield(want["zstar"]), ofield(want["psi_delta"])
    tol = ztol(Nx * Ny * Nz)
    # e_a: field parity, non-negative to roundoff, and its integral (:61-76)
    ea = ocb.Field(APE.AvailablePotentialEnergy(m, z)).compute().interior_ijk()
    ea_want = K.evaluate(K.local_ape_ccc, og, "ccc", opsi, ob, ozs)
    np.testing.assert_allclose(ea, ea_want, rtol=0, atol=1e-9 * np.abs(ea_want).max())
    assert ea.min() > -np.sqrt(np.finfo(np.float64).eps) * np.abs(vals).max()
    assert _integral(APE.AvailablePotentialEnergy(m, z)) == pytest.approx(K.integral(ea_want, og, "ccc"), rel=1e-9)
    # Υ = z✶ - z with zero volume mean (:541-563)
    ups_f = ocb.Field(APE.BuoyancyDisplacementPotential(m, z)).compute()
    zc = og.node(2, "c", og.indices("ccc")[2])
    np.testing.assert_allclose(ups_f.interior_ijk(), z.interior_ijk() - zc, rtol=0, atol=1e-15)
    np.testing.assert_allclose(ups_f.interior_ijk(), want["zstar"] - zc, rtol=0, atol=tol)
    assert abs(ocb.Field(ocb.Average(APE.BuoyancyDisplacementPotential(m, z))).compute().value()) <= np.sqrt(np.finfo(np.float64).eps)
    # ε_a handed b for Υ is χ / 2 to roundoff (:565-587); with its own Υ it matches the oracle
    chi = ocb.Field(TV.TracerVarianceDissipationRate(m, "b")).compute().interior_ijk()
    half = ocb.Field(APE.AvailablePotentialEnergyDissipationRate(m, z, upsilon=m.tracers.b)).compute().interior_ijk()
    np.testing.assert_allclose(half, chi / 2, rtol=1e-13, atol=1e-18)
    eps_a = ocb.Field(APE.AvailablePotentialEnergyDissipationRate(m, z)).compute().interior_ijk()
    oups = ofield(K.evaluate(K.upsilon_ccc, og, "ccc", ozs))
    eps_want = K.evaluate(K.ape_dissipation_rate_ccc, og, "ccc", oups, oc, None, None, ob)
    np.testing.assert_allclose(eps_a, eps_want, rtol=0, atol=1e-9 * np.abs(eps_want).max())
    # one sweep for the whole family
    grp = ocb.FusedDiagnostics(ocb.Integral(APE.AvailablePotentialEnergy(m, z)), ocb.Integral(APE.BackgroundPotentialEnergy(m, z)),
                               ocb.Integral(APE.AvailablePotentialEnergyDissipationRate(m, z, upsilon=ups_f))).compute()
    assert grp.plan.n_launches == 2
    assert grp.integrals()[2] == pytest.approx(K.integral(eps_want, og, "ccc"), rel=1e-8)


def test_constructor_errors():
    og, pg, ob, pb = _pair((4, 4, 4), (-1, 0), np.zeros((4, 4, 4)))
    m = _model(pg, pb)
    with pytest.raises(TypeError):
        APE.AvailablePotentialEnergy(m, zstar=pb)
    with pytest.raises(ValueError):
        APE.AvailablePotentialEnergyDissipationRate(m)         # no closure: no diffusive flux
    with pytest.raises(ValueError):
        APE.AvailablePotentialEnergy(m, location=("f", "c", "c"))
