#!/usr/bin/env python
"""bench.py -- Gcells/s of the fused budget sweep (BASELINE.json metric) on N B200s of one node.

A "step" is one pass of the hot path over the synthetic state: ONE fused sweep that emits the KE budget
{K, ADV, PR, uᵢbᵢ, ε} and the tracer-variance budget {χ, 2c∇·F} as 7 volume integrals (config 3, variant A:
5 input fields read once, 40 B/cell algorithmic).  N > 1: the domain is slab-decomposed along y, every step
exchanges the y halos of u, v, w, b, p over NCCL and all-reduces the integrals.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--size NX NY NZ] [--fields] [--impl reference]

Prints ONE JSON line (see the driver contract): value = whole-job Gcells/s with inputs resident in HBM;
e2e = the same through the public API with HOST (pinned) inputs streamed in every step; roofline = the sweep
kernel's algorithmic bytes / CUDA-event time against the measured HBM peak; cpu_baseline = the oracle's
C/OpenMP port timed on the host cores on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Gcells/s, fused KE + tracer-variance budget sweep (Float64)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--size", type=int, nargs=3, default=None, help="per-job grid (default 1024^3)")
    ap.add_argument("--fields", action="store_true", help="variant B: also materialise the 7 output fields (96 B/cell)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-overlap", action="store_true", help="N>1: exchange halos before the sweep instead of behind its interior rows")
    ap.add_argument("--transport", default="peer", choices=["peer", "nccl"],
                    help="N>1 halo exchange: direct stores into the neighbours' halos over NVLink (CUDA IPC), or NCCL send/recv")
    ap.add_argument("--cpu-size", type=int, default=256, help="edge of the CPU-baseline sample cube")
    return ap.parse_args()


# ---- clocks sampling during the timed region ---------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons of one GPU, polled on a thread during the timed region.  NVML is initialised when the
    sampler is CREATED (before the warm-up steps), not when it is entered: nvmlInit in all eight processes of a box at the
    start of the timed region stalled kernel launches for 15 - 40 ms (a 9.2 ms step read as 11.2 ms over 20 steps)."""

    def __init__(self, index=0, enabled=True):
        self.samples, self.reasons, self.max_mhz, self._stop = [], set(), None, threading.Event()
        self.index, self.nv, self.handle = index, None, None
        if enabled:
            try:
                import pynvml as nv
                nv.nvmlInit()
                self.nv, self.handle = nv, nv.nvmlDeviceGetHandleByIndex(index)
                self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(self.handle, nv.NVML_CLOCK_SM)
            except Exception as e:  # pragma: no cover
                self.error = repr(e)
        self.thread = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        try:
            nv, h = self.nv, self.handle
            if nv is None:
                return
            names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                     nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                     nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                     nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
            while not self._stop.is_set():
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
                time.sleep(0.02)
        except Exception as e:  # pragma: no cover
            self.error = repr(e)

    def __enter__(self):
        self.thread.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self.thread.join(timeout=2)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ---- synthetic state on the device ---------------------------------------------------------------------------
def make_state(ocb, torch, size, device, rank=0, world=1, seed=20261019):
    """SURVEY.md section 8d synthetic fields for this rank's y-slab; halos filled by the library."""
    Nx, Ny, Nz = size
    ny_local = Ny // world
    grid = ocb.RectilinearGrid((Nx, ny_local, Nz), x=(0, 1), y=(rank / world, (rank + 1) / world), z=(-1, 0),
                               topology=(ocb.Periodic, ocb.Periodic, ocb.Bounded), dtype=torch.float64, device=device)
    m = ocb.NonhydrostaticModel(grid, tracers=("b",), closure=ocb.Closure(nu=1e-6, kappa=1e-7))
    gen = torch.Generator(device=device)
    two_pi = 6.283185307179586
    x_f = torch.arange(Nx, device=device, dtype=torch.float64) / Nx
    x_c = x_f + 0.5 / Nx
    y_f = rank / world + torch.arange(ny_local, device=device, dtype=torch.float64) / Ny
    y_c = y_f + 0.5 / Ny
    z_c = -1 + (torch.arange(Nz, device=device, dtype=torch.float64) + 0.5) / Nz

    def noise(field, idx, amp):
        gen.manual_seed(seed + 17 * idx + 1000 * rank)
        it = field.interior
        for k0 in range(0, it.shape[0], 64):       # plane batches: bounded temporaries at 1024^3
            blk = it[k0:k0 + 64]
            blk.add_(torch.randn(blk.shape, generator=gen, device=device, dtype=torch.float64), alpha=amp)

    u, v, w = m.velocities.u, m.velocities.v, m.velocities.w
    b, p = m.tracers.b, m.pressures.pNHS
    u.interior.copy_((torch.sin(two_pi * x_f)[None, None, :] * torch.cos(two_pi * y_c)[None, :, None]).expand_as(u.interior))
    v.interior.copy_((-torch.cos(two_pi * x_c)[None, None, :] * torch.sin(two_pi * y_f)[None, :, None]).expand_as(v.interior))
    b.interior.copy_((1e-5 * z_c)[:, None, None].expand_as(b.interior))
    noise(u, 0, 0.1); noise(v, 1, 0.1); noise(w, 2, 0.1); noise(b, 3, 1e-6); noise(p, 4, 0.1)
    w.interior[0].zero_(); w.interior[-1].zero_()
    if world == 1:
        m.fill_halo_regions()
    return grid, m


def budget_ops(ocb, m):
    KE, TV = ocb.KineticEnergyEquation, ocb.TracerVarianceEquation
    return [KE.KineticEnergy(m), KE.KineticEnergyAdvection(m), KE.KineticEnergyPressureRedistribution(m),
            KE.PotentialEnergyConversion(m), KE.DissipationRate(m), TV.TracerVarianceDissipationRate(m, "b"),
            TV.TracerVarianceDiffusion(m, "b")]


def main():
    args = parse()
    if args.impl == "reference":
        from bench_reference import run_reference
        return run_reference(args)
    import torch
    import ocb_loader
    ocb = ocb_loader.load()

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        # NCCL's internal stream at high priority: its send / recv blocks are placed ahead of the sweep's pending blocks
        # whenever an SM frees up, so the ring exchange stays hidden behind the interior rows
        opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
        dist.init_process_group("nccl", device_id=device, pg_options=opts)
    size = tuple(args.size) if args.size else (1024, 1024, 1024)
    # weak scaling: every rank owns a full `size` slab stack => the job is size[0] x (size[1]*world) x size[2]
    job = (size[0], size[1] * world, size[2])
    grid, m = make_state(ocb, torch, job, device, rank, world)
    exchanger = None
    if world > 1:
        from oceanostics_b200.distributed import SlabHaloExchanger
        exchanger = SlabHaloExchanger(grid, [m.velocities.u, m.velocities.v, m.velocities.w, m.tracers.b, m.pressures.pNHS],
                                      depth=2, rank=rank, world=world)
        exchanger.exchange()
        for f in (m.velocities.u, m.velocities.v, m.velocities.w, m.tracers.b, m.pressures.pNHS):
            f.fill_halo_regions_xz()

    ops = budget_ops(ocb, m)
    members = [ocb.Integral(o) for o in ops]
    if args.fields:
        members += [ocb.Field(o) for o in ops]
    grp = ocb.FusedDiagnostics(*members)
    plan = grp.plan
    cells_local = grid.N[0] * grid.N[1] * grid.N[2]
    cells_job = cells_local * world
    overlapped = None
    if world > 1 and not args.fields and not args.no_overlap:
        from oceanostics_b200.distributed import OverlappedSlabBudget
        overlapped = OverlappedSlabBudget(grid, [m.velocities.u, m.velocities.v, m.velocities.w, m.tracers.b, m.pressures.pNHS],
                                          lambda: budget_ops(ocb, m), depth=2, rank=rank, world=world, transport=args.transport)
    elif world > 1 and args.transport == "peer":
        from oceanostics_b200.distributed import PeerSlabExchanger
        exchanger = PeerSlabExchanger(grid, [m.velocities.u, m.velocities.v, m.velocities.w, m.tracers.b, m.pressures.pNHS],
                                      depth=2, rank=rank, world=world)

    def step():
        if overlapped is not None:
            overlapped.step()
            return
        if exchanger is not None:
            exchanger.exchange()
        plan.execute()
        if dist is not None:
            dist.all_reduce(plan.integrals)

    sampler = ClockSampler(local, enabled=(rank == 0))      # NVML set up here, outside the timed region; rank 0 reports its board
    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    # ---- timed region: device-resident inputs ---------------------------------------------------------------
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    with sampler as clocks:
        torch.cuda.synchronize()
        ev0.record()
        for s in range(args.steps):
            if overlapped is not None:
                overlapped.step()
                continue
            if exchanger is not None:
                exchanger.exchange()
            kev[s][0].record()
            plan.execute()
            kev[s][1].record()
            if dist is not None:
                dist.all_reduce(plan.integrals)
        ev1.record()
        torch.cuda.synchronize()
    total_ms = ev0.elapsed_time(ev1)
    timeline = None
    if overlapped is not None:       # one more (untimed) step with event marks: where the exchange ends relative to the sweep
        tl = {}
        overlapped.step()
        overlapped.step(timeline=tl)
        torch.cuda.synchronize()
        timeline = overlapped.timeline_ms(tl)
    if overlapped is None:
        sweep_ms = sum(a.elapsed_time(b) for a, b in kev) / args.steps
    else:
        # the sweep kernel alone, timed after the region (the overlapped step interleaves three sweeps with the exchange)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(args.steps):
            plan.execute()
        b.record()
        torch.cuda.synchronize()
        sweep_ms = a.elapsed_time(b) / args.steps
    t = torch.tensor([total_ms, sweep_ms], dtype=torch.float64, device=device)
    per_rank = None
    if dist is not None:
        # every rank's own numbers (loop time per step, stand-alone sweep, where its marked step's phases end): the boards of a
        # box do not run at the same clock under their power caps, and the all-reduce makes every step as long as the slowest
        mine = [total_ms / args.steps, sweep_ms] + ([timeline[k] for k in ("exchange_end", "interior_end", "strips_end", "end")] if timeline else [0.0] * 4)
        every = [torch.zeros(6, dtype=torch.float64, device=device) for _ in range(world)]
        dist.all_gather(every, torch.tensor(mine, dtype=torch.float64, device=device))
        per_rank = [[round(float(x), 3) for x in e] for e in every]
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.barrier()
    total_ms, sweep_ms = float(t[0]), float(t[1])
    ms_per_step = total_ms / args.steps
    value = cells_job / (ms_per_step * 1e-3) / 1e9

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = plan.algorithmic_bytes / (sweep_ms * 1e-3) / 1e9
    # DRAM traffic of one launch of the sweep kernel, from the committed ncu captures of this exact workload
    # (dram__bytes_read.sum + dram__bytes_write.sum; class-sum kernel: profiles/r1_launches_bench1024_sums_v2.csv, mean of
    # five launches, captured with 36-column tiles -- the 34-column tiles adopted afterwards fetch 5.6 % less per tile and
    # were not re-captured; with output fields: profiles/r1_ncu_budget1024_fields_v13_metrics.csv); null for other shapes
    traffic = None
    if size == (1024, 1024, 1024) and plan.variant == 1:
        traffic = (47783394000 + 60439114000) if args.fields else (48463504384 + 1897165)
    integrals = grp.integrals() if overlapped is None else [float(x) for x in overlapped.total.cpu()]

    out = {
        "metric": METRIC, "value": round(value, 3), "unit": "Gcells/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": round(ms_per_step, 4), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"config 3{'B' if args.fields else 'A'}: {job[0]}x{job[1]}x{job[2]} Float64, fused KE budget "
                               "{K, ADV(Centered-2), PR, u_i b_i, eps} + tracer variance {chi, 2c div F}, 7 volume integrals"
                               + (" + 7 output fields" if args.fields else ""),
                   "per_gpu_grid": list(size), "halo": 3, "topology": "Periodic,Periodic,Bounded",
                   "parallelism": f"y-slabs x{world}" if world > 1 else "single GPU",
                   "l2": "inputs (5 x %.2f GB per GPU) exceed the 126 MB L2; no flush needed" % (cells_local * 8 / 1e9),
                   "kernel_variant": {0: "generic", 1: "fused-budget", 2: "fused-budget + generic"}.get(plan.variant, str(plan.variant))},
        "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                     "traffic": traffic, "kernel_ms": round(sweep_ms, 4), "algorithmic_bytes_per_cell": plan.algorithmic_bytes / cells_local,
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)"},
        "gpu_launches": ((3 if overlapped is not None else 1) * plan.n_launches + ((3 if args.transport == "peer" else 2) if exchanger is not None else 0)) * args.steps,   # + push, signal, wait | pack_many, unpack_many
        "clocks": clocks.summary(),
        "integrals": integrals,
    }
    if world > 1:
        out["comm"] = {"step_minus_sweep_ms": round(ms_per_step - sweep_ms, 4), "overlapped": overlapped is not None, "timeline_ms": timeline,
                       "per_rank_ms": {"columns": ["loop_step", "sweep_alone", "exchange_end", "interior_end", "strips_end", "end"], "rows": per_rank},
                       "halo_bytes_per_rank_per_step": exchanger.bytes_per_exchange,
                       "transport": args.transport,
                       "note": "exchange (peer: one kernel storing the edge rows into the neighbours' halos over NVLink + flag words; nccl: "
                               "pack, grouped send/recv ring, unpack) + all-reduce of the 7 integrals; when overlapped it runs on a side "
                               "stream behind the interior rows and only the edge strips wait for it"}
        px = overlapped.exchanger if overlapped is not None else exchanger
        if hasattr(px, "check"):
            px.check()                                   # a halo wait that timed out invalidates the run
    # ---- end to end with HOST inputs: every rank streams its own slab from pinned host memory -----------------------
    e2e = None
    if not args.no_e2e:
        import psutil
        need = 5 * cells_local * 8 * 1.05 * world            # all ranks of this node pin their slab's five inputs
        ok = psutil.virtual_memory().available > 1.4 * need
        flag = torch.tensor([1.0 if ok else 0.0], device=device)
        if dist is not None:
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if float(flag) > 0:
            from bench_e2e import run_e2e
            e2e = run_e2e(ocb, torch, grid, m, args, dist=dist, exchanger=exchanger)
            t = torch.tensor([e2e["ms_per_step"]], dtype=torch.float64, device=device)
            if dist is not None:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e["ms_per_step"] = round(float(t[0]), 3)
            e2e["value"] = round(cells_job / (float(t[0]) * 1e-3) / 1e9, 4)
            e2e["h2d_bytes_per_step"] *= world
            e2e["d2h_bytes_per_step"] *= world
        else:
            e2e = {"value": None, "unit": "Gcells/s", "h2d_bytes_per_step": None, "d2h_bytes_per_step": None,
                   "skipped": "not enough free host memory to pin every rank's inputs"}
    if rank == 0:
        if e2e is not None:
            out["e2e"] = e2e
        if not args.no_cpu and world == 1:
            from bench_reference import cpu_baseline
            out["cpu_baseline"] = cpu_baseline(args)
        print(json.dumps(out))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
