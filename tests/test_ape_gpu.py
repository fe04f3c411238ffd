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


def _stride_scrambled(values, shape):
    N = values.size
    scr = np.zeros(N)
    for m in range(N):
        scr[(7 * m) % N] = values[m]                     # a bijection when gcd(7, N) = 1
    return scr.reshape(shape, order="F")


def test_vertical_sort_column_known_profile():
    """:280-298 and :405-441: the sorted column on its 1 x 1 x N grid recovers the tanh profile, integrates like the model
    grid, keeps the topology, and BackgroundPotentialEnergy / AvailablePotentialEnergy accept it."""
    Nx, Ny, Nz = 4, 3, 5
    N, Lz = Nx * Ny * Nz, 1.0
    zs = -Lz + (np.arange(1, N + 1) - 0.5) * Lz / N
    bs = 0.1 * np.tanh((zs + Lz / 2) / 0.2)
    og, pg, ob, pb = _pair((Nx, Ny, Nz), (-Lz, 0), _stride_scrambled(bs, (Nx, Ny, Nz)))
    m = _model(pg, pb)
    col = APE.reference_height(m, method=BPE.VerticalSort())
    torch.cuda.synchronize()
    assert col.interior_ijk().shape == (1, 1, N) and col.grid.user_topology == pg.user_topology
    want = OB.vertical_sort(ob)
    np.testing.assert_allclose(col.interior_ijk().ravel(), zs, rtol=1e-13)
    np.testing.assert_array_equal(col.reference_buoyancy.interior_ijk().ravel(), bs)
    np.testing.assert_allclose(col.sorted_height.interior_ijk().ravel(), want["sorted_height"], rtol=0, atol=1e-15)
    np.testing.assert_allclose(col.reference_potential.interior_ijk().ravel(), want["psi_delta"], rtol=1e-12, atol=1e-16)
    ranked = APE.reference_height(m)
    eb_col, eb_rank = _integral(APE.BackgroundPotentialEnergy(m, col)), _integral(APE.BackgroundPotentialEnergy(m, ranked))
    assert eb_col == pytest.approx(float(np.sum(-bs * zs) * (1.0 / N)), rel=1e-12) and eb_col == pytest.approx(eb_rank, rel=1e-12)
    assert _integral(APE.AvailablePotentialEnergy(m, col)) == pytest.approx(_integral(APE.AvailablePotentialEnergy(m, ranked)), rel=1e-10)
    with pytest.raises(ValueError):
        APE.BuoyancyDisplacementPotential(m, col)          # Υ needs z✶ on the model grid


def test_profile_lookup_three_ways_matches_the_ranked_sort():
    """:226-275: ProfileLookup(), ProfileLookup(column), ProfileLookup(b✶, z✶) reproduce ThreeDimensionalSort on a tie-free
    field; malformed profiles are rejected with the reference's messages."""
    Nx, Ny, Nz = 3, 2, 4
    N = Nx * Ny * Nz
    og, pg, ob, pb = _pair((Nx, Ny, Nz), (-1, 0), _stride_scrambled(-0.5 + np.arange(N) / (N - 1), (Nx, Ny, Nz)))
    m = _model(pg, pb)
    ranked = APE.reference_height(m)
    eb, ea = _integral(APE.BackgroundPotentialEnergy(m, ranked)), _integral(APE.AvailablePotentialEnergy(m, ranked))
    col = APE.reference_height(m, method=BPE.VerticalSort())
    bstar = col.reference_buoyancy.interior.reshape(-1).clone()
    zprof = col.interior.reshape(-1).clone()
    want = OB.profile_lookup(ob, bstar.cpu().numpy(), zprof.cpu().numpy())
    for method in (BPE.ProfileLookup(), BPE.ProfileLookup(col), BPE.ProfileLookup(bstar, zprof)):
        z = APE.reference_height(m, method=method)
        torch.cuda.synchronize()
        np.testing.assert_allclose(z.interior_ijk(), ranked.interior_ijk(), rtol=1e-13)
        np.testing.assert_allclose(z.interior_ijk(), want["zstar"], rtol=1e-13)
        np.testing.assert_allclose(z.reference_potential.interior_ijk(), want["psi_delta"], rtol=1e-11, atol=1e-15)
        assert _integral(APE.BackgroundPotentialEnergy(m, z)) == pytest.approx(eb, rel=1e-12)
        assert _integral(APE.AvailablePotentialEnergy(m, z)) == pytest.approx(ea, rel=1e-10)
        e = ocb.Field(APE.AvailablePotentialEnergy(m, z)).compute().interior_ijk()
        assert e.min() > -np.sqrt(np.finfo(np.float64).eps) * max(np.abs(e).max(), np.finfo(np.float64).eps)
    for msg, bad in (("buoyancy and height have different", BPE.ProfileLookup(bstar[:-1], zprof)),
                     ("ordered from the densest fluid up", BPE.ProfileLookup(bstar.flip(0), zprof)),
                     ("heights paired with", BPE.ProfileLookup(bstar, zprof.flip(0)))):
        with pytest.raises(ValueError, match=msg):
            APE.reference_height(m, method=bad)
    with pytest.raises(TypeError):
        APE.reference_height(m, method=BPE.ProfileLookup(3.0))


def test_profile_lookup_with_ties_against_the_oracle():
    """Tied buoyancy classes: every cell of a class takes the mid-height of the class's run of slots (:556-571)."""
    rng = np.random.default_rng(23)
    Nx, Ny, Nz = 12, 10, 8
    vals = np.round(rng.standard_normal((Nx, Ny, Nz)) * 3) / 3          # many repeated levels
    og, pg, ob, pb = _pair((Nx, Ny, Nz), (-1, 0), vals)
    m = _model(pg, pb)
    z = APE.reference_height(m, method=BPE.ProfileLookup())
    torch.cuda.synchronize()
    col = OB.vertical_sort(ob)
    want = OB.profile_lookup(ob, col["b_sorted"], col["zstar"])
    np.testing.assert_allclose(z.interior_ijk(), want["zstar"], rtol=0, atol=ztol(vals.size))
    heav = APE.reference_height(m, method=BPE.HeavisideIntegral())       # the same band mid-heights, by construction
    np.testing.assert_allclose(z.interior_ijk(), heav.interior_ijk(), rtol=0, atol=ztol(vals.size))
