"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: the slab partition, the ring halo-exchange schedule and
the all-reduce of the integral vector.  The pack / unpack kernels are CUDA-only, so the CPU test injects
torch-slicing packers with the same buffer layout; the CUDA packers are tested in test_multigpu_gpu.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import ocb


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _pack_torch(field, side, h, buf):
    Hy, ext = field.halo[1], field.stored[1]
    j0 = Hy if side == 0 else Hy + ext - h
    buf.copy_(field.data[:, j0:j0 + h, :])


def _unpack_torch(field, side, h, buf):
    Hy, ext = field.halo[1], field.stored[1]
    j0 = Hy - h if side == 0 else Hy + ext
    field.data[:, j0:j0 + h, :].copy_(buf)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oceanostics_b200.distributed import SlabHaloExchanger, slab_range, allreduce_integrals
    Nx, Ny, Nz, H = 6, 4 * world, 4, 3
    lo, hi = slab_range(Ny, rank, world)
    g = ocb.RectilinearGrid((Nx, hi - lo + 1, Nz), extent=(1, 1, 1), device="cpu")
    f = ocb.Field(("c", "c", "c"), g)
    w = ocb.Field(("c", "c", "f"), g)
    # value = global linear index of the cell, so halos can be checked against the periodic global field
    for fld in (f, w):
        nz = fld.stored[2]
        kk, jj, ii = torch.meshgrid(torch.arange(nz), torch.arange(lo - 1, hi), torch.arange(Nx), indexing="ij")
        fld.interior.copy_((kk * 1000 + jj * 10 + ii).double())
    ex = SlabHaloExchanger(g, [f, w], depth=2, pack=_pack_torch, unpack=_unpack_torch)
    ex.exchange()
    ok = True
    for fld in (f, w):
        nz = fld.stored[2]
        for m in range(1, 3):   # two halo planes each side
            below = ((lo - 1 - m) % Ny)
            above = ((hi - 1 + m) % Ny)
            got_lo = fld.data[H:H + nz, H - m, H:H + Nx]
            got_hi = fld.data[H:H + nz, H + (hi - lo + 1) - 1 + m, H:H + Nx]
            kk, ii = torch.meshgrid(torch.arange(nz), torch.arange(Nx), indexing="ij")
            ok &= bool(torch.equal(got_lo, (kk * 1000 + below * 10 + ii).double()))
            ok &= bool(torch.equal(got_hi, (kk * 1000 + above * 10 + ii).double()))
    ints = torch.tensor([float(rank + 1), 10.0 * (rank + 1)], dtype=torch.float64)
    allreduce_integrals(ints)
    tot = world * (world + 1) / 2
    ok &= ints.tolist() == [tot, 10.0 * tot]
    out[rank] = ok
    dist.destroy_process_group()


def test_slab_ranges_partition_the_axis():
    from oceanostics_b200.distributed import slab_range
    for n, world in ((1024, 8), (10, 4), (7, 2)):
        r = [slab_range(n, k, world) for k in range(world)]
        assert r[0][0] == 1 and r[-1][1] == n
        assert all(r[k][1] + 1 == r[k + 1][0] for k in range(world - 1))


@pytest.mark.timeout(120)
def test_ring_halo_exchange_and_allreduce_world2():
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    assert out[0] and out[1]


@pytest.mark.timeout(180)
def test_ring_halo_exchange_and_allreduce_world4():
    """Four slabs: every rank has two DISTINCT neighbours (at world 2 both sides are the same peer)."""
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(4, _free_port(), out), nprocs=4, join=True)
    assert all(out[r] for r in range(4))


def _filter_worker(rank, world, port, out, periodic):
    """The slab filter's row exchange (the filter kernels themselves are CUDA-only): every rank's extended copy of ψ
    must hold the global rows lo-w .. hi+w (wrapped when y is periodic, absent past a true boundary)."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oceanostics_b200.distributed import SlabFilteredField, slab_range
    Nx, Ny, Nz = 6, 5 * world + 1, 4
    lo, hi = slab_range(Ny, rank, world)
    topo = (ocb.Periodic, ocb.Periodic if periodic else ocb.Bounded, ocb.Bounded)
    g = ocb.RectilinearGrid((Nx, hi - lo + 1, Nz), x=(0, 1), y=((lo - 1) / Ny, hi / Ny), z=(-1, 0), topology=topo, device="cpu")
    f = ocb.Field(("c", "c", "c"), g)
    kk, jj, ii = torch.meshgrid(torch.arange(Nz), torch.arange(lo - 1, hi), torch.arange(Nx), indexing="ij")
    f.interior.copy_((kk * 1000 + jj * 10 + ii).double())
    sf = SlabFilteredField(f, lambda psi: ocb.SpatialFilters.BoxFilter(psi, dims=(1, 2), N=7), rank=rank, world=world, periodic_y=periodic)
    sf.exchange()
    w = 3
    assert sf.w == w
    rows = list(range(lo - 1 - sf.ext_lo, hi + sf.ext_hi))            # 0-based global rows of the extended slab
    assert sf.ext_lo == (w if periodic or rank > 0 else 0) and sf.ext_hi == (w if periodic or rank < world - 1 else 0)
    kk, jj, ii = torch.meshgrid(torch.arange(Nz), torch.tensor([r % Ny for r in rows]), torch.arange(Nx), indexing="ij")
    out[rank] = bool(torch.equal(sf.ext_psi.interior, (kk * 1000 + jj * 10 + ii).double()))
    dist.destroy_process_group()


@pytest.mark.timeout(240)
@pytest.mark.parametrize("world,periodic", [(2, True), (4, True), (3, False)])
def test_slab_filter_row_exchange(world, periodic):
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_filter_worker, args=(world, _free_port(), out, periodic), nprocs=world, join=True)
    assert all(out[r] for r in range(world))
