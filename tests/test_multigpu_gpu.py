"""Slab-decomposed evaluation on the device: the CUDA pack / unpack kernels, the x/z halo fill of a slab, and the
equality of per-slab integrals (summed) with the single-domain integrals.  With one GPU the two ranks are emulated as
two slabs on the same device exchanging their packed buffers directly (the exchange schedule itself is covered by the
world-size-2 gloo test); with >= 2 GPUs the NCCL path runs under torchrun in bench.py."""
import ctypes as C

import numpy as np
import pytest
import torch

from helpers import State, ocb, K
from test_fused_budget_gpu import budget

pytestmark = pytest.mark.gpu


def _slab_model(st, lo, hi):
    """A model on rows lo..hi (1-based, inclusive) of the global state, y halos left unset."""
    Nx, Ny, Nz = st.pg.N
    ny = hi - lo + 1
    g = ocb.RectilinearGrid((Nx, ny, Nz), x=(0, 1), y=((lo - 1) / Ny, hi / Ny), z=(-1, 0), device=st.pg.device)
    m = ocb.NonhydrostaticModel(g, tracers=("b",), closure=st.closures()["scalar"][1])
    H = 3
    for dst, src in ((m.velocities.u, st.u), (m.velocities.v, st.v), (m.velocities.w, st.w), (m.tracers.b, st.b), (m.pressures.pNHS, st.p)):
        dst.interior.copy_(src.interior[:, lo - 1:hi, :])
    return g, m


def test_two_slabs_reproduce_the_global_integrals():
    from oceanostics_b200.distributed import _pack_cuda, _unpack_cuda
    st = State(size=(64, 48, 24), device="cuda")
    ops, _ = budget(st)
    ref = ocb.FusedDiagnostics(*[ocb.Integral(o) for o in ops]).compute().integrals()
    slabs = [_slab_model(st, 1, 24), _slab_model(st, 25, 48)]
    h = 2
    flds = [[m.velocities.u, m.velocities.v, m.velocities.w, m.tracers.b, m.pressures.pNHS] for _, m in slabs]
    for a, b in zip(flds[0], flds[1]):       # ring of two: each slab's low and high neighbour is the other slab
        bufs = {}
        for name, f in (("a", a), ("b", b)):
            for side in (0, 1):
                buf = torch.empty((f.data.shape[0], h, f.data.shape[2]), dtype=torch.float64, device="cuda")
                _pack_cuda(f, side, h, buf)
                bufs[(name, side)] = buf
        _unpack_cuda(a, 0, h, bufs[("b", 1)]); _unpack_cuda(a, 1, h, bufs[("b", 0)])
        _unpack_cuda(b, 0, h, bufs[("a", 1)]); _unpack_cuda(b, 1, h, bufs[("a", 0)])
        a.fill_halo_regions_xz(); b.fill_halo_regions_xz()
    total = np.zeros(len(ops))
    for (g, m), _ in zip(slabs, range(2)):
        KE, TV = ocb.KineticEnergyEquation, ocb.TracerVarianceEquation
        sops = [KE.KineticEnergy(m), KE.KineticEnergyAdvection(m), KE.KineticEnergyPressureRedistribution(m),
                KE.PotentialEnergyConversion(m), KE.DissipationRate(m), TV.TracerVarianceDissipationRate(m, "b"),
                TV.TracerVarianceDiffusion(m, "b")]
        grp = ocb.FusedDiagnostics(*[ocb.Integral(o) for o in sops]).compute()
        total += np.array(grp.integrals())
    for o, a, b in zip(ops, total, ref):
        scale = float(ocb.Field(o).compute().interior.abs().sum()) / (64 * 48 * 24)
        assert abs(a - b) <= 1e-12 * scale, o.name


def test_self_exchange_equals_periodic_fill():
    from oceanostics_b200.distributed import SlabHaloExchanger
    st = State(size=(32, 16, 8), device="cuda")
    f = ocb.Field(("c", "c", "c"), st.pg)
    f.data.copy_(st.b.data)
    want = st.b.data.clone()
    H = 3
    f.data[:, :H, :] = -1.0
    f.data[:, -H:, :] = -1.0
    SlabHaloExchanger(st.pg, [f], depth=3, rank=0, world=1).exchange()
    torch.cuda.synchronize()
    assert torch.equal(f.data[H:-H], want[H:-H])


def test_peer_self_exchange_equals_periodic_fill():
    """A ring of one rank through the peer-memory path (push kernel + flag words + stream-ordered wait): the slab's own
    edge rows land in its own halos, i.e. the periodic fill; repeated steps count the flags up and never time out."""
    from oceanostics_b200.distributed import PeerSlabExchanger
    st = State(size=(32, 16, 8), device="cuda")
    fields = []
    for src in (st.u, st.w, st.b):
        f = ocb.Field(src.location, st.pg)
        f.data.copy_(src.data)
        f.data[:, :3, :] = -1.0
        f.data[:, -3:, :] = -1.0
        fields.append(f)
    px = PeerSlabExchanger(st.pg, fields, depth=3, rank=0, world=1, timeout=2.0)
    for _ in range(3):
        px.exchange()
    torch.cuda.synchronize()
    px.check()
    assert px.step_no == 3 and px.flags[:3].tolist() == [3, 3, 0]
    for f, src in zip(fields, (st.u, st.w, st.b)):
        assert torch.equal(f.data[3:-3], src.data[3:-3])
    # a wait for a step nobody pushed gives up after the timeout and reports it
    px.step_no += 1
    px.timeout = 0.05
    px.wait()
    torch.cuda.synchronize()
    with pytest.raises(ocb.OcbError, match="did not arrive"):
        px.check()


@pytest.mark.parametrize("size", [(64, 48, 24), (70, 21, 9), (64, 33, 40), (32, 16, 5)], ids=str)
def test_ring_mode_of_the_class_sum_kernel(size):
    """Whole periodic y extent: the class-sum kernel gives every row the interior item weights and drops the extra row
    jy_hi + 1 (its items are the wrapped row 1's).  Same integrals as the window form (OCB_SUMS_NO_RING=1) and the oracle."""
    import os
    from test_fused_budget_gpu import oracle_budget
    st = State(size=size, device="cuda")
    ops, oc = budget(st)
    want = oracle_budget(st, oc)
    ring = ocb.FusedDiagnostics(*[ocb.Integral(o) for o in ops]).compute()
    os.environ["OCB_SUMS_NO_RING"] = "1"
    try:
        plain = ocb.FusedDiagnostics(*[ocb.Integral(o) for o in ops]).compute()
    finally:
        os.environ.pop("OCB_SUMS_NO_RING", None)
    assert ring.plan.variant == 1 and ring.plan.ring_capable and not plain.plan.ring_capable
    V = 1.0 / (size[0] * size[1] * size[2])
    for o, w_, a, b in zip(ops, want, ring.integrals(), plain.integrals()):
        ref, scale = float(np.sum(w_)) * V, float(np.sum(np.abs(w_))) * V
        assert abs(a - ref) <= 1e-12 * scale, (o.name, a, ref)
        assert abs(a - b) <= 1e-12 * scale, (o.name, a, b)


def test_ring_step_on_a_ring_of_one():
    """RingSlabBudget with one rank: the push lands the slab's own edge rows in its own halos, the sweep's edge tile rows
    wait for the flag words in the kernel, the reduction kernel passes through the (one-rank) mailboxes.  Wiped y halos must
    be refilled by the step itself; repeated steps count the flags up; the integrals are those of the periodic domain."""
    from oceanostics_b200.distributed import RingSlabBudget
    st = State(size=(64, 80, 24), device="cuda")
    ops, _ = budget(st)
    ref = ocb.FusedDiagnostics(*[ocb.Integral(o) for o in ops]).compute().integrals()
    scales = [float(ocb.Field(o).compute().interior.abs().sum()) / (64 * 80 * 24) for o in ops]
    fields = [st.u, st.v, st.w, st.b, st.p]
    rb = RingSlabBudget(st.pg, fields, lambda: budget(st)[0], rank=0, world=1, timeout=5.0)
    assert not rb.serial
    for f in fields:
        f.data[:, :3, :] = 0.0
        f.data[:, -3:, :] = 0.0
    for _ in range(3):
        got = rb.step().clone()
    torch.cuda.synchronize()
    rb.check()
    assert rb.ring.block[:3].tolist() == [3, 3, 0]
    for o, a, b, sc in zip(ops, got.tolist(), ref, scales):
        assert abs(a - b) <= 1e-12 * sc, (o.name, a, b)
    # a thin slab (fewer than three tile rows) runs the push on the compute stream first; same numbers
    st2 = State(size=(64, 24, 12), device="cuda")
    ops2, _ = budget(st2)
    ref2 = ocb.FusedDiagnostics(*[ocb.Integral(o) for o in ops2]).compute().integrals()
    rb2 = RingSlabBudget(st2.pg, [st2.u, st2.v, st2.w, st2.b, st2.p], lambda: budget(st2)[0], rank=0, world=1, timeout=5.0)
    assert rb2.serial
    got2 = rb2.step().tolist()
    scales2 = [float(ocb.Field(o).compute().interior.abs().sum()) / (64 * 24 * 12) for o in ops2]
    for o, a, b, sc in zip(ops2, got2, ref2, scales2):
        assert abs(a - b) <= 1e-12 * sc, (o.name, a, b)
    # a step whose push never happened: the gated tiles give up after the timeout, the integrals read NaN, the word is set
    rb.ring.desc.timeout_seconds = 0.05
    rb.exchanger.step_no += 1
    rb.plan.execute_ring(rb.ring, rb.exchanger.step_no)
    torch.cuda.synchronize()
    assert bool(torch.isnan(rb.total).all())
    with pytest.raises(ocb.OcbError, match="did not arrive"):
        rb.check()


def test_operands_on_a_non_current_device_are_guarded():
    """Every entry point runs on the device that owns its operands (OcbDeviceGuard), whatever the current device is."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    st = State(size=(32, 16, 8), device="cuda:1")
    ops, _ = budget(st)
    with torch.cuda.device(0):
        got = ocb.FusedDiagnostics(*[ocb.Integral(o) for o in ops]).compute().integrals()
    with torch.cuda.device(1):
        want = ocb.FusedDiagnostics(*[ocb.Integral(o) for o in ops]).compute().integrals()
    assert got == want


def test_batched_pack_equals_the_per_field_kernels():
    """ocb_halo_pack_y_many (both edges of all fields in one launch) fills the same buffers, and writes the same halo rows,
    as ocb_halo_pack_y / ocb_halo_unpack_y called field by field (u is Face in x, w has Nz + 1 levels: different shapes)."""
    from oceanostics_b200.distributed import _pack_cuda, _unpack_cuda, _pack_many_cuda
    st = State(size=(32, 16, 8), device="cuda")
    fields = [st.u, st.v, st.w, st.b, st.p]
    h = 2
    mk = lambda f: torch.zeros((f.data.shape[0], h, f.data.shape[2]), dtype=torch.float64, device="cuda")
    one = [[mk(f), mk(f)] for f in fields]
    many = [[mk(f), mk(f)] for f in fields]
    for f, (lo, hi) in zip(fields, one):
        _pack_cuda(f, 0, h, lo); _pack_cuda(f, 1, h, hi)
    _pack_many_cuda(fields, h, [b for pair in many for b in pair], unpack=False)
    torch.cuda.synchronize()
    for a, b in zip(one, many):
        assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and float(a[0].abs().sum()) > 0
    # unpack: swap the buffers (a ring of one) and compare the halo rows written by the two paths
    ref = []
    for f, (lo, hi) in zip(fields, one):
        g = f.data.clone()
        _unpack_cuda(f, 0, h, hi); _unpack_cuda(f, 1, h, lo)
        ref.append(f.data.clone())
        f.data.copy_(g)
    _pack_many_cuda(fields, h, [b for lo, hi in many for b in (hi, lo)], unpack=True)
    torch.cuda.synchronize()
    for f, want in zip(fields, ref):
        assert torch.equal(f.data, want)


def test_host_streamed_budget_matches_resident_sweep():
    from oceanostics_b200.streaming import HostStreamedBudget
    st = State(size=(64, 40, 36), device="cuda")
    ops, _ = budget(st)
    ref = ocb.FusedDiagnostics(*[ocb.Integral(o) for o in ops]).compute().integrals()
    m_fields = [st.u, st.v, st.w, st.b, st.p]
    hs = HostStreamedBudget(st.pg, m_fields, ops, n_slabs=5)
    hs.load_host_from_device()
    for f in m_fields:
        f.data.zero_()                    # the device copies must really come from the host mirrors
    hs.step()
    got = hs.integrals()
    assert all(g.plan.variant == 1 for g in hs.groups)
    for o, a, b in zip(ops, got, ref):
        scale = float(ocb.Field(o).compute().interior.abs().sum()) / (64 * 40 * 36)
        assert abs(a - b) <= 1e-12 * scale, o.name
    assert hs.h2d_bytes == sum(f.data.numel() * 8 for f in m_fields)


@pytest.mark.parametrize("kind", ["gauss", "box"])
def test_slab_filter_self_exchange_equals_periodic_filter(kind):
    """One slab spanning the whole periodic y axis: the rows it borrows are its own, so the slab filter must reproduce
    the single-domain filter (same taps, weights and order for every retained output)."""
    from oceanostics_b200.distributed import SlabFilteredField
    st = State(size=(40, 24, 18), device="cuda")
    SF = ocb.SpatialFilters
    mk = (lambda f: SF.GaussianFilter(f, dims=(1, 2, 3), sigma=2.0 / 24, N=9)) if kind == "gauss" else \
         (lambda f: SF.BoxFilter(f, dims=(2, 3), N=7))
    got = SlabFilteredField(st.b, mk, rank=0, world=1).compute()
    want = ocb.Field(mk(st.b)).compute()
    torch.cuda.synchronize()
    np.testing.assert_allclose(got.interior.cpu().numpy(), want.interior.cpu().numpy(), rtol=1e-13, atol=1e-15)
    # a filter that does not touch y needs no exchange at all
    only_x = SlabFilteredField(st.b, lambda f: SF.BoxFilter(f, dims=(1,), N=5), rank=0, world=1)
    assert only_x.w == 0 and only_x.bytes_per_exchange == 0
    np.testing.assert_array_equal(only_x.compute().interior.cpu().numpy(),
                                  ocb.Field(SF.BoxFilter(st.b, dims=(1,), N=5)).compute().interior.cpu().numpy())


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
@pytest.mark.parametrize("stretched", [False, True], ids=["regular", "stretched"])
def test_nccl_slabs_match_single_domain(stretched):
    """torchrun x2: NCCL halo exchange + all-reduced integrals equal the single-domain integrals; so do the overlapped, the
    peer-memory and the ring steps (on a stretched z axis: through the per-plane-metric kernel variants)."""
    import os
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(here, "multigpu_check.py")]
    env = dict(os.environ)
    if stretched:
        env["MG_STRETCHED"] = "1"
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600, env=env)
    assert r.returncode == 0 and "MULTIGPU_CHECK OK" in r.stdout, r.stdout[-3000:]
