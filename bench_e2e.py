"""bench.py's `e2e` object: the same metric with HOST (pinned) inputs -- every step copies all five input fields
host -> device, runs the fused sweep slab by slab behind the copies, and reads the seven integrals back."""
from __future__ import annotations


def run_e2e(ocb, torch, grid, m, args, dist=None, exchanger=None):
    from oceanostics_b200.streaming import HostStreamedBudget
    from bench import budget_ops
    fields = [m.velocities.u, m.velocities.v, m.velocities.w, m.tracers.b, m.pressures.pNHS]
    hs = HostStreamedBudget(grid, fields, budget_ops(ocb, m), n_slabs=16)
    hs.load_host_from_device()
    steps = max(2, min(args.steps, 5))
    def step():
        # N > 1: the host mirrors hold each rank's slab with the y halos of the last exchange already in them (an
        # Oceananigans Distributed model hands over filled halos), so the streamed step needs only the all-reduce
        hs.step()
        if dist is not None:
            dist.all_reduce(hs.total)
            hs.result_host.copy_(hs.total, non_blocking=True)

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    cells = grid.N[0] * grid.N[1] * grid.N[2]
    return {"value": round(cells / (ms * 1e-3) / 1e9, 4), "unit": "Gcells/s", "h2d_bytes_per_step": hs.h2d_bytes,
            "d2h_bytes_per_step": hs.d2h_bytes, "ms_per_step": round(ms, 3), "steps": steps,
            "path": "HostStreamedBudget: pinned host parents -> 16 z slabs, copy stream overlapped with the fused sweeps, "
                    "integrals summed on device and read back", "integrals": hs.integrals()}
