"""GPU parity tests of the sort-based background potential energy (K3): the hand-written radix sort alone, and the
reference height / Ψδ / E_b against the NumPy oracle on the reference's own analytic cases
(test/test_ape_diagnostics.jl:85-107 two-cell column, :405-461 scrambled tanh profile with ties, :162-191 method
agreement, :71 volume-mean height) -- comparing only permutation-invariant quantities, as the reference does."""
import ctypes as C

import numpy as np
import pytest
import torch

from helpers import ocb, mirror
from oracle.grid import OGrid, OField
from oracle import bpe as OB

pytestmark = pytest.mark.gpu
BPE = ocb.BackgroundPotentialEnergyEquation


def sort_pairs(keys: torch.Tensor):
    n = keys.numel()
    out = torch.empty_like(keys)
    perm = torch.empty(n, dtype=torch.int32, device=keys.device)
    dt = ocb._lib.OCB_F64 if keys.dtype == torch.float64 else ocb._lib.OCB_F32
    ocb._lib.check(ocb._lib.lib().ocb_sort_pairs(dt, keys.data_ptr(), out.data_ptr(), perm.data_ptr(), n, None))
    torch.cuda.synchronize()
    return out, perm.long() & 0xFFFFFFFF


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("n", [1, 31, 4096, 4097, 100_003, 3_000_000])
def test_radix_sort_is_sorted_stable_and_a_permutation(n, dtype):
    g = torch.Generator(device="cuda").manual_seed(n)
    keys = torch.randn(n, generator=g, device="cuda", dtype=dtype)
    keys[::7] = keys[0].clone()              # many ties
    if n > 8:
        keys[3], keys[5] = 0.0, -0.0         # isless: -0.0 sorts before +0.0
        keys[1] = -float("inf"); keys[2] = float("inf")
    out, perm = sort_pairs(keys)
    kn, pn = keys.cpu().numpy(), perm.cpu().numpy()
    want = np.argsort(OB.order_key(kn), kind="stable")      # the reference's CPU order: isless + stable
    np.testing.assert_array_equal(pn, want)
    np.testing.assert_array_equal(out.cpu().numpy(), kn[want])


def test_chained_scatter_stress():
    """The chained scatter waits on other tiles' flags and reads their counts behind relaxed loads: many sorts of awkward
    sizes and digit distributions (all keys equal, two values, few distinct digits, sorted and reversed input, group-boundary
    tile counts 15/16/17 and 255/256/257 tiles), each checked for order, stability (through the permutation) and content."""
    g = torch.Generator(device="cuda").manual_seed(20261020)
    tile = 4096
    sizes = [tile * t + d for t in (15, 16, 17, 255, 256, 257) for d in (-1, 0, 1)] + [7, 300_001, 1_234_567, 2_500_000]
    for trial, n in enumerate(sizes * 2):
        kind = trial % 6
        if kind == 0:
            keys = torch.randn(n, generator=g, device="cuda", dtype=torch.float64)
        elif kind == 1:
            keys = torch.full((n,), 1.25, device="cuda", dtype=torch.float64)
        elif kind == 2:
            keys = (torch.rand(n, generator=g, device="cuda") < 0.5).double() - 0.5
        elif kind == 3:
            keys = torch.randint(0, 7, (n,), generator=g, device="cuda").double() * 2.0 ** -20
        elif kind == 4:
            keys = torch.arange(n, device="cuda", dtype=torch.float64)
        else:
            keys = -torch.arange(n, device="cuda", dtype=torch.float64) * 1e-3
        out, perm = sort_pairs(keys)
        assert bool((out[1:] >= out[:-1]).all()), (n, kind)
        assert bool((keys[perm] == out).all()), (n, kind)
        same = out[1:] == out[:-1]
        assert bool((perm[1:][same] > perm[:-1][same]).all()), (n, kind)          # stable: ties keep their input order
        assert int(perm.sum()) == n * (n - 1) // 2, (n, kind)                      # a permutation


def ztol(n, Lz=1.0):
    """z✶ = z_bot + (cumulative volume)/A: the reference accumulates the N sorted volumes sequentially (cumsum!), the
    library with a tree scan; both carry up to ~N eps/2 of rounding relative to the column depth, so heights are
    compared to max(1e-12, 4 N eps) * Lz (absolute), not relative to z✶ itself (which passes through 0 at the top)."""
    return max(1e-12, 4 * n * np.finfo(np.float64).eps) * Lz


def _pair(size, zspec, b_values, topology=("P", "P", "B")):
    topo = {"P": ocb.Periodic, "B": ocb.Bounded, "F": ocb.Flat}
    og = OGrid(size, x=(0, 1), y=(0, 1), z=zspec, topology=topology)
    pg = ocb.RectilinearGrid(size, x=(0, 1), y=(0, 1), z=zspec, topology=tuple(topo[t] for t in topology), device="cuda")
    ob = OField(og, "ccc")
    ob.interior[...] = b_values
    ob.fill_halo()
    return og, pg, ob, mirror(ob, pg)


def test_two_cell_column_analytic():
    """test/test_ape_diagnostics.jl:85-107: light water under dense water in a two-cell column."""
    og, pg, ob, pb = _pair((1, 1, 2), (-1, 0), np.array([1.0, 0.0]).reshape(1, 1, 2))
    z = BPE.reference_height(pb)
    torch.cuda.synchronize()
    np.testing.assert_allclose(z.interior_ijk().ravel(), [-0.25, -0.75], rtol=1e-14)
    m = ocb.NonhydrostaticModel(pg)
    m.tracers.b = pb
    eb = ocb.Field(ocb.Integral(BPE.BackgroundPotentialEnergy(m, z))).compute().value()
    assert abs(eb - (-(1.0 * -0.25) * 0.5)) < 1e-14          # ∫E_b = -b z✶ ΔV summed = 0.125 ... dense cell at z✶=-0.75 has b=0
    want = OB.reference_height(ob)
    np.testing.assert_allclose(z.reference_potential.interior_ijk(), want["psi_delta"], rtol=1e-13, atol=1e-15)


@pytest.mark.parametrize("method", ["ThreeDimensionalSort", "HeavisideIntegral"])
def test_scrambled_tanh_profile_with_ties(method):
    """test/test_ape_diagnostics.jl:405-461: a horizontally uniform tanh profile, scrambled in z (so whole planes tie)."""
    Nx, Ny, Nz = 8, 6, 16
    rng = np.random.default_rng(5)
    zc = -1 + (np.arange(Nz) + 0.5) / Nz
    prof = np.tanh(4 * (zc + 0.5))
    scr = prof[rng.permutation(Nz)]
    og, pg, ob, pb = _pair((Nx, Ny, Nz), (-1, 0), np.broadcast_to(scr, (Nx, Ny, Nz)))
    z = BPE.reference_height(pb, method=getattr(BPE, method)())
    torch.cuda.synchronize()
    want = OB.reference_height(ob, method)
    got_z = z.interior_ijk()
    # permutation-invariant comparisons: sorted b*, the multiset of z*, the volume-mean height, ∫E_b
    np.testing.assert_array_equal(z.sorted_buoyancy.cpu().numpy(), want["b_sorted"])
    np.testing.assert_allclose(np.sort(got_z.ravel()), np.sort(want["zstar"].ravel()), rtol=0, atol=ztol(got_z.size))
    dV = want["dV"]
    assert abs(np.sum(got_z * dV) - np.sum(og.node(2, "c", og.indices("ccc")[2]) * dV)) < 1e-12       # :71
    eb_got, eb_want = np.sum(-ob.interior * got_z * dV), np.sum(-ob.interior * want["zstar"] * dV)
    assert abs(eb_got - eb_want) <= 1e-12 * abs(eb_want)
    if method == "HeavisideIntegral":        # tie-free semantics: every cell of a tied plane gets the same height
        np.testing.assert_allclose(got_z, want["zstar"], rtol=0, atol=ztol(got_z.size))
        np.testing.assert_allclose(z.reference_potential.interior_ijk(), want["psi_delta"], rtol=1e-10, atol=1e-13)
    else:                                    # stable sort == Base.sortperm! on the CPU: even the permutation agrees
        np.testing.assert_array_equal(z.permutation.cpu().numpy() - 1, want["perm"])
        np.testing.assert_allclose(got_z, want["zstar"], rtol=0, atol=ztol(got_z.size))
        np.testing.assert_allclose(z.reference_potential.interior_ijk(), want["psi_delta"], rtol=1e-10, atol=1e-13)


def test_random_field_methods_agree_and_stretched_grid():
    """:162-191 method agreement on a tie-free field; :368-390 only HeavisideIntegral takes a stretched grid."""
    from helpers import stretched_faces
    rng = np.random.default_rng(11)
    Nx, Ny, Nz = 24, 20, 18
    vals = rng.standard_normal((Nx, Ny, Nz)) * 1e-6 + 1e-5 * (-1 + (np.arange(Nz) + 0.5) / Nz)
    og, pg, ob, pb = _pair((Nx, Ny, Nz), (-1, 0), vals)
    z3 = BPE.reference_height(pb)
    zh = BPE.reference_height(pb, method=BPE.HeavisideIntegral())
    torch.cuda.synchronize()
    tol = ztol(Nx * Ny * Nz)
    np.testing.assert_allclose(z3.interior_ijk(), zh.interior_ijk(), rtol=0, atol=tol)
    want = OB.reference_height(ob)
    np.testing.assert_allclose(z3.interior_ijk(), want["zstar"], rtol=0, atol=tol)
    np.testing.assert_allclose(z3.reference_potential.interior_ijk(), want["psi_delta"], rtol=1e-9, atol=1e-18)
    zf = stretched_faces(Nz)
    og, pg, ob, pb = _pair((Nx, Ny, Nz), zf, vals)
    with pytest.raises(ValueError):
        BPE.reference_height(pb)
    zs = BPE.reference_height(pb, method=BPE.HeavisideIntegral())
    torch.cuda.synchronize()
    want = OB.reference_height(ob, "HeavisideIntegral")
    np.testing.assert_allclose(zs.interior_ijk(), want["zstar"], rtol=0, atol=tol)


def test_bucketed_scatter_equals_the_direct_one(monkeypatch):
    """Above 2^20 cells the ThreeDimensionalSort epilogue buckets the (cell, rank) pairs by destination region before the final
    scatter; forced here at a size the oracle reaches: the same z*, Psi_delta, permutation and sorted buoyancy bit for bit."""
    rng = np.random.default_rng(21)
    Nx, Ny, Nz = 40, 33, 29
    vals = rng.standard_normal((Nx, Ny, Nz)) * 1e-6 + 1e-5 * (-1 + (np.arange(Nz) + 0.5) / Nz)
    vals[3:9, :, 7] = vals[3, 0, 7]
    og, pg, ob, pb = _pair((Nx, Ny, Nz), (-1, 0), vals)
    direct = BPE.reference_height(pb)
    torch.cuda.synchronize()
    monkeypatch.setenv("OCB_BPE_BUCKETED", "1")
    pb2 = mirror(ob, pg)
    buck = BPE.reference_height(pb2)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(buck.interior_ijk(), direct.interior_ijk())
    np.testing.assert_array_equal(buck.reference_potential.interior_ijk(), direct.reference_potential.interior_ijk())
    np.testing.assert_array_equal(buck.permutation.cpu().numpy(), direct.permutation.cpu().numpy())
    np.testing.assert_array_equal(buck.sorted_buoyancy.cpu().numpy(), direct.sorted_buoyancy.cpu().numpy())
    want = OB.reference_height(ob)
    np.testing.assert_allclose(buck.interior_ijk(), want["zstar"], rtol=0, atol=ztol(Nx * Ny * Nz))


def test_recompute_on_evolution_and_unknown_methods():
    rng = np.random.default_rng(2)
    og, pg, ob, pb = _pair((8, 8, 8), (-1, 0), rng.standard_normal((8, 8, 8)))
    z = BPE.reference_height(pb)
    before = z.interior_ijk().copy()
    pb.interior.mul_(-1.0)                        # flipping the sign reverses the order
    z.compute(time=1.0)
    torch.cuda.synchronize()
    after = z.interior_ijk()
    np.testing.assert_allclose(after, -1.0 - before, rtol=0, atol=ztol(512))       # slot centres mirror about the mid-depth
    with pytest.raises(TypeError):
        BPE.reference_height(pb, method="ThreeDimensionalSort")          # a method object, not a name


def test_full_size_config5_reference_height_properties():
    """BASELINE.json configs[4] at its full size (N = 67 108 864 keys) through properties that need no oracle: the sorted
    buoyancy is non-decreasing and a rearrangement of the input (same sum and extremes), the permutation is one (every
    index once), z✶ conserves the volume-mean height, the two methods agree on a tie-free field, and e_a >= 0."""
    free, _ = torch.cuda.mem_get_info()
    if free < 20e9:
        pytest.skip("needs ~15 GB of device memory")
    Nx, Ny, Nz = 512, 512, 256
    N = Nx * Ny * Nz
    g = ocb.RectilinearGrid((Nx, Ny, Nz), extent=(1, 1, 0.5), device="cuda")
    m = ocb.NonhydrostaticModel(g, tracers=("b",))
    b = m.tracers.b
    gen = torch.Generator(device="cuda").manual_seed(11)
    zc = -0.5 + (torch.arange(Nz, device="cuda", dtype=torch.float64) + 0.5) * 0.5 / Nz
    b.interior.copy_((1e-5 * zc)[:, None, None] + 1e-6 * torch.randn(b.interior.shape, generator=gen, device="cuda", dtype=torch.float64))
    b.fill_halo_regions()
    z3 = BPE.reference_height(m)
    bs = z3.sorted_buoyancy
    assert bool(torch.all(bs[1:] >= bs[:-1]))
    assert float(bs[0]) == float(b.interior.min()) and float(bs[-1]) == float(b.interior.max())
    assert abs(float(bs.sum()) - float(b.interior.sum())) <= 1e-12 * float(b.interior.abs().sum())
    perm = z3.permutation
    seen = torch.zeros(N, dtype=torch.int32, device="cuda")
    seen.index_add_(0, perm - 1, torch.ones(N, dtype=torch.int32, device="cuda"))
    assert int(seen.min()) == 1 and int(seen.max()) == 1
    tol = ztol(N, 0.5)
    assert abs(float(z3.interior.mean()) - (-0.25)) <= tol                     # volume-mean z✶ = volume-mean z (uniform cells)
    zh = BPE.reference_height(m, method=BPE.HeavisideIntegral())
    assert float((z3.interior - zh.interior).abs().max()) <= tol
    APE = ocb.AvailablePotentialEnergyEquation
    ea = ocb.Field(APE.AvailablePotentialEnergy(m, z3)).compute().interior
    assert float(ea.min()) > -np.sqrt(np.finfo(np.float64).eps) * float(b.interior.abs().max())
