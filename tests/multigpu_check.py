"""Run under torchrun on >= 2 GPUs (tests/test_multigpu_gpu.py::test_nccl_slabs_match_single_domain launches it):
every rank builds the same global synthetic state, keeps its y slab, exchanges halos over NCCL, runs the fused budget
sweep and all-reduces the integrals; rank 0 compares with the single-domain integrals."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ocb_loader  # noqa: E402

ocb = ocb_loader.load()


def budget_ops(m):
    KE, TV = ocb.KineticEnergyEquation, ocb.TracerVarianceEquation
    return [KE.KineticEnergy(m), KE.KineticEnergyAdvection(m), KE.KineticEnergyPressureRedistribution(m),
            KE.PotentialEnergyConversion(m), KE.DissipationRate(m), TV.TracerVarianceDissipationRate(m, "b"),
            TV.TracerVarianceDiffusion(m, "b")]


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from oceanostics_b200.distributed import SlabHaloExchanger, slab_range, allreduce_integrals
    Nx, Ny, Nz = 128, 192, 32          # slabs of 96 / 48 rows at 2 / 4 ranks (gated ring step), 24 rows at 8 (serial ring step)
    gen = torch.Generator(device=dev).manual_seed(1234)
    glob = {n: torch.randn((Nz + (1 if n == "w" else 0), Ny, Nx), generator=gen, device=dev, dtype=torch.float64) for n in "uvwbp"}
    glob["w"][0].zero_(); glob["w"][-1].zero_()

    # MG_STRETCHED=1: a tanh-stretched z axis -- the same steps then run on the per-plane-metric variants of the kernels
    zspec = (-1, 0)
    if os.environ.get("MG_STRETCHED"):
        kf = np.arange(Nz + 1)
        zspec = -np.tanh(2.0 * (1 - kf / Nz)) / np.tanh(2.0)

    def model_on(lo, hi, full):
        g = ocb.RectilinearGrid((Nx, hi - lo + 1, Nz), x=(0, 1), y=((lo - 1) / Ny, hi / Ny), z=zspec, device=dev)
        m = ocb.NonhydrostaticModel(g, tracers=("b",), closure=ocb.Closure(nu=1e-3, kappa=1e-4))
        flds = {"u": m.velocities.u, "v": m.velocities.v, "w": m.velocities.w, "b": m.tracers.b, "p": m.pressures.pNHS}
        for n, f in flds.items():
            f.interior.copy_(glob[n][:, lo - 1:hi, :])
        if full:
            m.fill_halo_regions()
        return g, m, list(flds.values())

    lo, hi = slab_range(Ny, rank, world)
    g, m, flds = model_on(lo, hi, full=False)
    ex = SlabHaloExchanger(g, flds, depth=2, rank=rank, world=world)
    ex.exchange()
    for f in flds:
        f.fill_halo_regions_xz()
    grp = ocb.FusedDiagnostics(*[ocb.Integral(o) for o in budget_ops(m)])
    grp.compute()
    total = grp.plan.integrals.clone()
    allreduce_integrals(total)
    from oceanostics_b200.distributed import OverlappedSlabBudget
    ov = OverlappedSlabBudget(g, flds, lambda: budget_ops(m), depth=2, rank=rank, world=world)
    for f in flds:                       # wipe the y halos: the overlapped step must refill them itself
        f.data[:, :3, :] = 0.0
        f.data[:, -3:, :] = 0.0
    ov_total = ov.step().clone()
    # ---- the same step with the direct peer-memory exchange (CUDA IPC stores into the neighbours' halos) ---------------
    from oceanostics_b200.distributed import PeerSlabExchanger
    before = [f.data.clone() for f in flds]              # halos as NCCL left them
    for f in flds:
        f.data[:, :3, :] = 0.0
        f.data[:, -3:, :] = 0.0
    torch.cuda.synchronize(); dist.barrier()
    px = PeerSlabExchanger(g, flds, depth=2, rank=rank, world=world)
    px.exchange()
    torch.cuda.synchronize(); dist.barrier()
    px.check()
    peer_same = all(torch.equal(f.data[:, 1:3, :], b[:, 1:3, :]) and torch.equal(f.data[:, -3:-1, :], b[:, -3:-1, :]) for f, b in zip(flds, before))
    ovp = OverlappedSlabBudget(g, flds, lambda: budget_ops(m), depth=2, rank=rank, world=world, transport="peer")
    for f in flds:
        f.data[:, :3, :] = 0.0
        f.data[:, -3:, :] = 0.0
    torch.cuda.synchronize(); dist.barrier()
    ovp_total = None
    for _ in range(3):                                   # repeated steps: the flag words count up, the all-reduce orders re-use
        ovp_total = ovp.step().clone()
    torch.cuda.synchronize()
    ovp.exchanger.check()
    # ---- the ring step: one-row push, gated class-sum sweep, in-kernel all-reduce over the ranks' sync blocks -------------
    from oceanostics_b200.distributed import RingSlabBudget
    rb = RingSlabBudget(g, flds, lambda: budget_ops(m), rank=rank, world=world, timeout=20.0)
    for f in flds:
        f.data[:, :3, :] = 0.0
        f.data[:, -3:, :] = 0.0
    torch.cuda.synchronize(); dist.barrier()
    ring_total = None
    for _ in range(4):
        ring_total = rb.step().clone()
    torch.cuda.synchronize()
    rb.check()
    same = ring_total.clone()
    dist.all_reduce(same, op=dist.ReduceOp.MAX)
    ring_same_bits = torch.tensor([1.0 if bool((same == ring_total).all()) else 0.0], device=dev)
    dist.all_reduce(ring_same_bits, op=dist.ReduceOp.MIN)
    p_ok = torch.tensor([1.0 if peer_same else 0.0], device=dev)
    dist.all_reduce(p_ok, op=dist.ReduceOp.MIN)
    # ---- slab filter: the y pass reaches 5 rows into each neighbour (exchanged over NCCL) -------------------------
    from oceanostics_b200.distributed import SlabFilteredField
    SF = ocb.SpatialFilters
    sigma = 2.5 / Ny
    mk = lambda f: SF.GaussianFilter(f, dims=(1, 2, 3), sigma=sigma, N=11)
    slab_f = SlabFilteredField(m.tracers.b, mk, rank=rank, world=world).compute()
    _, mg_all, _ = model_on(1, Ny, full=True)
    want_f = ocb.Field(mk(mg_all.tracers.b)).compute().interior[:, lo - 1:hi, :]
    f_err = float((slab_f.interior - want_f).abs().max()) / float(want_f.abs().max())
    f_ok = torch.tensor([1.0 if f_err <= 1e-14 else 0.0], device=dev)
    dist.all_reduce(f_ok, op=dist.ReduceOp.MIN)
    ok = bool(f_ok.item() > 0) and bool(p_ok.item() > 0) and bool(ring_same_bits.item() > 0)
    if rank == 0:
        print("slab filter max rel err", f_err, flush=True)
        _, mg, _ = model_on(1, Ny, full=True)
        ref = ocb.FusedDiagnostics(*[ocb.Integral(o) for o in budget_ops(mg)]).compute()
        want = ref.plan.integrals.cpu().numpy()
        got = total.cpu().numpy()
        scales = [float(ocb.Field(o).compute().interior.abs().sum()) / (Nx * Ny * Nz) for o in budget_ops(mg)]
        for a, b, s in zip(got, want, scales):
            ok &= abs(a - b) <= 1e-12 * s
        for a, b, s in zip(ov_total.cpu().numpy(), want, scales):
            ok &= abs(a - b) <= 1e-12 * s
        for a, b, s in zip(ovp_total.cpu().numpy(), want, scales):
            ok &= abs(a - b) <= 1e-12 * s
        for a, b, s in zip(ring_total.cpu().numpy(), want, scales):
            ok &= abs(a - b) <= 1e-12 * s
        print("ring step integrals:", ring_total.tolist(), "identical on every rank:", bool(ring_same_bits.item() > 0), flush=True)
        print("peer exchange equals NCCL exchange:", bool(p_ok.item() > 0), flush=True)
        print("MULTIGPU_CHECK", "OK" if ok else "FAIL", got.tolist(), want.tolist(), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
