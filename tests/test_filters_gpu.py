"""GPU parity tests of the separable spatial filters (K2, through ocb_filter_execute) against the NumPy oracle, which
is itself pinned to the reference's brute-force references and golden vectors (test_oracle_golden.py); plus the
reference's own golden vectors re-checked directly on the device result (test/test_spatial_filters.jl:381-431,
658-745, 695-712, 210-222)."""
import numpy as np
import pytest
import torch

from helpers import ocb, mirror, make_grids, stretched_faces
from oracle.grid import OGrid, OField
from oracle import filters as OF

pytestmark = pytest.mark.gpu
SF = ocb.SpatialFilters


def pair(size=(8, 8, 8), topology=("P", "P", "P"), halo=(2, 2, 2), loc="ccc", fn=None, seed=3, z=None, dtype=np.float64):
    """Oracle field and its device mirror on the same grid."""
    topo = {"P": ocb.Periodic, "B": ocb.Bounded, "F": ocb.Flat}
    tdt = torch.float64 if dtype == np.float64 else torch.float32
    zz = z if z is not None else (0, 1)
    og = OGrid(size, x=(0, 1), y=(0, 1), z=zz, halo=halo, topology=topology, dtype=dtype)
    pg = ocb.RectilinearGrid(size, x=(0, 1), y=(0, 1), z=zz, halo=halo, topology=tuple(topo[t] for t in topology), dtype=tdt, device="cuda")
    of = OField(og, loc)
    if fn is None:
        of.interior[...] = np.random.default_rng(seed).standard_normal(of.interior.shape).astype(dtype)
    else:
        of.set(fn)
    of.fill_halo(impenetrable=False)
    return og, pg, of, mirror(of, pg)


def run(op):
    f = ocb.Field(op).compute()
    torch.cuda.synchronize()
    return f.interior_ijk()


@pytest.mark.parametrize("dims", [(1,), (2,), (3,), (1, 2), (1, 3), (1, 2, 3)])
@pytest.mark.parametrize("kind", ["box", "gauss"])
def test_periodic_filters(kind, dims):
    og, pg, of, pf = pair(size=(12, 10, 9))
    if kind == "box":
        got, want = run(SF.BoxFilter(pf, dims=dims, N=5)), OF.box_filter(of, dims, 5)
    else:
        got, want = run(SF.GaussianFilter(pf, dims=dims, sigma=1.3 / 8, N=7)), OF.gaussian_filter(of, dims, 1.3 / 8, 7)
    np.testing.assert_allclose(got, want, rtol=1e-13, atol=1e-15)


@pytest.mark.parametrize("boundary", ["shrink", "edge", dict(left=-1.0, right=2.5)], ids=["shrink", "edge", "constant"])
@pytest.mark.parametrize("kind", ["box", "gauss"])
@pytest.mark.parametrize("loc", ["ccc", "ccf", "fff"])
def test_bounded_policies_and_face_extents(kind, boundary, loc):
    og, pg, of, pf = pair(size=(9, 8, 7), topology=("B", "B", "B"), loc=loc)
    assert pf.extent == of.ext
    dims = (1, 2, 3)
    if kind == "box":
        got, want = run(SF.BoxFilter(pf, dims=dims, N=5, boundary=boundary)), OF.box_filter(of, dims, 5, boundary)
    else:
        got = run(SF.GaussianFilter(pf, dims=dims, sigma=0.2, N=(5, 7, 3), boundary=boundary))
        want = OF.gaussian_filter(of, dims, 0.2, (5, 7, 3), boundary)
    assert got.shape == of.ext
    np.testing.assert_allclose(got, want, rtol=1e-13, atol=1e-15)


def test_box_filter_hand_computed_golden_vectors():
    """test/test_spatial_filters.jl:381-431, checked on the device result itself."""
    g = ocb.RectilinearGrid(10, x=(0, 1), halo=2, topology=(ocb.Periodic, ocb.Flat, ocb.Flat), device="cuda")
    c = ocb.CenterField(g)
    c.interior[0, 0, :] = torch.arange(10.0, dtype=torch.float64)
    np.testing.assert_allclose(run(SF.BoxFilter(c, dims=(1,), N=3)).ravel(), [10 / 3, 1, 2, 3, 4, 5, 6, 7, 8, 17 / 3], rtol=1e-14)
    np.testing.assert_allclose(run(SF.BoxFilter(c, dims=(1,), N=5)).ravel(), [4, 3, 2, 3, 4, 5, 6, 7, 6, 5], rtol=1e-14)
    g = ocb.RectilinearGrid(10, x=(0, 1), halo=1, topology=(ocb.Bounded, ocb.Flat, ocb.Flat), device="cuda")
    c = ocb.CenterField(g)
    c.interior[0, 0, :] = torch.arange(10.0, dtype=torch.float64)
    chk = lambda N, exp, **kw: np.testing.assert_allclose(run(SF.BoxFilter(c, dims=(1,), N=N, **kw)).ravel(), exp, rtol=1e-14)
    chk(5, [1, 1.5, 2, 3, 4, 5, 6, 7, 7.5, 8])
    chk(7, [1.5, 2, 2.5, 3, 4, 5, 6, 6.5, 7, 7.5])
    chk(5, [0.6, 1.2, 2, 3, 4, 5, 6, 7, 7.8, 8.4], boundary="edge")
    chk(7, [6 / 7, 10 / 7, 15 / 7, 3, 4, 5, 6, 48 / 7, 53 / 7, 57 / 7], boundary="edge")
    chk(5, [0.2, 1, 2, 3, 4, 5, 6, 7, 5.6, 4], boundary=dict(left=-1.0, right=-2.0))
    chk(7, [3 / 7, 8 / 7, 2, 3, 4, 5, 6, 37 / 7, 31 / 7, 24 / 7], boundary=dict(left=-1.0, right=-2.0))


def test_gaussian_dirac_delta_impulse_response():
    """test/test_spatial_filters.jl:714-745."""
    Nn = 21
    g = ocb.RectilinearGrid(Nn, x=(0, 1), halo=3, topology=(ocb.Periodic, ocb.Flat, ocb.Flat), device="cuda")
    i0 = Nn // 2
    for width, sc in [(3, 1.0), (3, 2.0), (5, 1.5)]:
        c = ocb.CenterField(g)
        c.interior[0, 0, i0] = 1.0
        out = run(SF.GaussianFilter(c, dims=(1,), sigma=sc / Nn, N=2 * width + 1)).ravel()
        wts = [np.exp(-(i - width) ** 2 / (2 * sc ** 2)) for i in range(2 * width + 1)]
        exp = np.zeros(Nn)
        for j in range(Nn):
            if abs(i0 - j) <= width:
                exp[j] = wts[i0 - j + width] / sum(wts)
        np.testing.assert_allclose(out, exp, rtol=1e-12, atol=1e-300)


def test_gaussian_identity_on_face_field_bit_exact():
    """test/test_spatial_filters.jl:695-712: σ << Δ with N=3 returns the input bit for bit, last face included."""
    og, pg, of, pf = pair(topology=("P", "P", "B"), loc="ccf", fn=lambda x, y, z: np.sin(2 * np.pi * x) + z ** 2)
    for dims in ((3,), (1, 3), (1, 2, 3)):
        out = run(SF.GaussianFilter(pf, dims=dims, sigma=1e-4, N=3))
        assert np.all(np.isfinite(out)) and np.array_equal(out, of.interior)


@pytest.mark.parametrize("boundary", ["shrink", "edge"])
def test_stretched_gaussian(boundary):
    """StretchedGaussianFilterKernel (gaussian_filter.jl:193-262) on the reference's stretched z grid and on a
    stretched periodic axis."""
    N = 12
    zf = stretched_faces(N)
    og, pg, of, pf = pair(size=(8, 8, N), topology=("P", "P", "B"), halo=(3, 3, 3), z=zf)
    got = run(SF.GaussianFilter(pf, dims=(1, 3), sigma=0.15, boundary=boundary))
    want = OF.gaussian_filter(of, (1, 3), 0.15, None, boundary)
    np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-15)
    og, pg, of, pf = pair(size=(8, 8, N), topology=("P", "P", "B"), halo=(3, 3, 3), z=zf, loc="ccf")
    got = run(SF.GaussianFilter(pf, dims=(3,), sigma=0.1, N=5, boundary=boundary))
    np.testing.assert_allclose(got, OF.gaussian_filter(of, (3,), 0.1, 5, boundary), rtol=1e-12, atol=1e-15)


def test_width_inference_tuple_order_and_windowed_output():
    og, pg, of, pf = pair(fn=lambda x, y, z: np.sin(2 * np.pi * x) + np.cos(2 * np.pi * z))
    s = 1.5 / 8
    assert np.array_equal(run(SF.GaussianFilter(pf, dims=(1, 3), sigma=s)), run(SF.GaussianFilter(pf, dims=(1, 3), sigma=s, N=7)))
    assert np.array_equal(run(SF.GaussianFilter(pf, dims=(3, 1), sigma=s, N=(5, 7))), run(SF.GaussianFilter(pf, dims=(1, 3), sigma=s, N=(7, 5))))
    full = run(SF.BoxFilter(pf, dims=(1, 2, 3), N=3))
    win = ocb.Field(SF.BoxFilter(pf, dims=(1, 2, 3), N=3), indices=((2, 6), None, (4, 4)))       # test_spatial_filters.jl:210-222
    win.compute()
    np.testing.assert_array_equal(win.interior_ijk(), full[1:6, :, 3:4])


def test_filter_of_an_operation_and_operator_form():
    """filter(KFO): the operand is materialised once, then filtered; `GaussianFilter(; ...)` builds a callable."""
    from helpers import State, K
    st = State(size=(16, 12, 10), device="cuda")
    m = st.model(closure=st.closures()["scalar"][1])
    ke = ocb.KineticEnergyEquation.KineticEnergy(m)
    filt = SF.GaussianFilter(dims=(1, 2), sigma=0.1, N=5)
    got = run(filt(ke))
    oke = OField(st.og, "ccc")
    oke.interior[...] = K.evaluate(K.kinetic_energy_ccc, st.og, "ccc", st.ou, st.ov, st.ow)
    np.testing.assert_allclose(got, OF.gaussian_filter(oke, (1, 2), 0.1, 5), rtol=1e-12)


def test_float32_filter():
    og, pg, of, pf = pair(size=(16, 8, 8), dtype=np.float32)
    got = run(SF.GaussianFilter(pf, dims=(1, 2, 3), sigma=0.2, N=5))
    want = OF.gaussian_filter(of, (1, 2, 3), 0.2, 5)
    assert np.linalg.norm(got.astype(np.float64) - want) <= 1e-5 * np.linalg.norm(want)


def test_full_size_config5_filter_properties():
    """BASELINE.json configs[4] at its full size (512 x 512 x 256 Float64, 17-tap Gaussian, :shrink in z) through
    properties that need no oracle: a field of ones is returned bit for bit (every pass divides the same accumulated
    weight sum), exact power-of-two scaling, linearity to roundoff, and a horizontally periodic box filter conserves the
    sum over x and y of every plane."""
    free, _ = torch.cuda.mem_get_info()
    if free < 8e9:
        pytest.skip("needs ~6 GB of device memory")
    Nx, Ny, Nz = 512, 512, 256
    g = ocb.RectilinearGrid((Nx, Ny, Nz), extent=(1, 1, 0.5), device="cuda")
    a, b = ocb.CenterField(g), ocb.CenterField(g)
    gen = torch.Generator(device="cuda").manual_seed(5)
    a.interior.copy_(torch.randn(a.interior.shape, generator=gen, device="cuda", dtype=torch.float64))
    b.interior.fill_(1.0)
    filt = SF.GaussianFilter(dims=(1, 2, 3), sigma=4.0 / 512, boundary="shrink")
    ones = ocb.Field(filt(b)).compute()
    assert filt(b).widths == (8, 8, 8)
    assert torch.equal(ones.interior, torch.ones_like(ones.interior))
    fa = ocb.Field(filt(a)).compute().interior.clone()
    a.interior.mul_(2.0)
    assert torch.equal(ocb.Field(filt(a)).compute().interior, 2.0 * fa)
    a.interior.mul_(0.5).add_(b.interior)                     # a + 1
    lin = ocb.Field(filt(a)).compute().interior
    assert float((lin - (fa + 1.0)).abs().max()) <= 1e-13 * float(fa.abs().max() + 1.0)
    box = ocb.Field(SF.BoxFilter(a, dims=(1, 2), N=17)).compute().interior
    assert float((box.sum(dim=(1, 2)) - a.interior.sum(dim=(1, 2))).abs().max()) <= 1e-9 * float(a.interior.abs().sum(dim=(1, 2)).max())
