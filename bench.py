#!/usr/bin/env python
"""bench.py -- Gcells/s of the fused budget sweep (BASELINE.json metric) on N B200s of one node.

A "step" is one pass of the hot path over the synthetic state: ONE fused sweep that emits the KE budget
{K, ADV, PR, uᵢbᵢ, ε} and the tracer-variance budget {χ, 2c∇·F} as 7 volume integrals (config 3, variant A:
5 input fields read once, 40 B/cell algorithmic).  N > 1: the domain is slab-decomposed along y, every step
exchanges the y halos of u, v, w, b, p over NCCL and all-reduces the integrals.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--size NX NY NZ] [--fields] [--impl reference]

N > 1 is STRONG scaling (BASELINE.json configs[2] as written): the 1024^3 grid is ONE job cut into N y slabs; the
per-GPU-1024^3 run of round 1 is kept as the `weak` sub-object.  N = 1 adds `secondary`: the other BASELINE configs
(2, 3B, 4, 5), each timed with CUDA events against its own roofline.

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
    ap.add_argument("--transport", default="ring", choices=["ring", "peer", "nccl"],
                    help="N>1: ring = peer-memory push + gated sweep + in-kernel all-reduce (3 launches); peer = peer-memory push, "
                         "interior/strip sweeps, ncclAllReduce; nccl = pack + ncclSend/Recv + unpack, ncclAllReduce")
    ap.add_argument("--no-weak", action="store_true", help="N>1: skip the per-GPU-1024^3 (weak scaling) sub-run")
    ap.add_argument("--no-secondary", action="store_true", help="N=1: skip the other BASELINE configs (secondary object)")
    ap.add_argument("--cpu-size", type=int, default=None, help="edge of the CPU sample cube (cpu_baseline: 256, --impl reference: 512)")
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


def slab_inputs(m):
    return [m.velocities.u, m.velocities.v, m.velocities.w, m.tracers.b, m.pressures.pNHS]


def integrand_scales(ocb, torch, grid, m, dist, nplanes=8):
    """sum |term| V of every budget term, estimated from `nplanes` planes in mid-depth (written as fields by the budget
    kernel) and scaled to the whole depth: the yardstick of the N > 1 parity check (ADV and PR nearly cancel in the sum)."""
    Nz = grid.N[2]
    k0 = max(1, Nz // 2 - nplanes // 2)
    k1 = min(Nz, k0 + nplanes - 1)
    fs = [ocb.Field(o, indices=(None, None, (k0, k1))) for o in budget_ops(ocb, m)]
    ocb.FusedDiagnostics(*fs).compute()
    V = float(grid.spacing0[0]) * float(grid.spacing0[1]) * float(grid.spacing0[2])
    s = torch.stack([f.interior.abs().sum() for f in fs]) * (V * Nz / (k1 - k0 + 1))
    if dist is not None:
        dist.all_reduce(s)
    return s


class Runner:
    """One workload (a job grid cut into `world` y slabs) with its step, built once and timed by `timed_run`."""

    def __init__(self, ocb, torch, args, job, device, rank, world, dist):
        self.ocb, self.torch, self.args, self.dist, self.world, self.rank, self.device = ocb, torch, args, dist, world, rank, device
        self.grid, self.m = make_state(ocb, torch, job, device, rank, world)
        grid, m = self.grid, self.m
        self.fields = slab_inputs(m)
        if world > 1:
            from oceanostics_b200.distributed import SlabHaloExchanger
            self.nccl_exchanger = SlabHaloExchanger(grid, self.fields, depth=2, rank=rank, world=world)
            self.nccl_exchanger.exchange()
            for f in self.fields:
                f.fill_halo_regions_xz()
        ops = budget_ops(ocb, m)
        members = [ocb.Integral(o) for o in ops]
        if args.fields:
            members += [ocb.Field(o) for o in ops]
        self.grp = ocb.FusedDiagnostics(*members)
        self.plan = self.grp.plan
        self.cells_local = grid.N[0] * grid.N[1] * grid.N[2]
        self.cells_job = self.cells_local * world
        self.transport = args.transport if world > 1 else None
        self.ring = self.overlapped = self.exchanger = None
        if world > 1:
            from oceanostics_b200 import distributed as D
            if args.fields and self.transport == "ring":
                self.transport = "peer"          # the ring step is the integral-only class-sum sweep
            if self.transport == "ring":
                self.ring = D.RingSlabBudget(grid, self.fields, lambda: budget_ops(ocb, m), rank=rank, world=world)
            elif not args.fields and not args.no_overlap:
                self.overlapped = D.OverlappedSlabBudget(grid, self.fields, lambda: budget_ops(ocb, m), depth=2, rank=rank, world=world,
                                                         transport=self.transport)
            elif self.transport == "peer":
                self.exchanger = D.PeerSlabExchanger(grid, self.fields, depth=2, rank=rank, world=world)
            else:
                self.exchanger = self.nccl_exchanger

    def step(self):
        if self.ring is not None:
            return self.ring.step()
        if self.overlapped is not None:
            return self.overlapped.step()
        if self.exchanger is not None:
            self.exchanger.exchange()
        self.plan.execute()
        if self.dist is not None:
            self.dist.all_reduce(self.plan.integrals)
        return self.plan.integrals

    def launches_per_step(self):
        if self.ring is not None:
            return 3                                                 # push (+ flag), gated sweep, partials + ring all-reduce
        n = self.plan.n_launches
        if self.overlapped is not None:
            return 3 * n + (2 if self.transport == "peer" else 2)    # + push, wait | pack_many, unpack_many
        if self.exchanger is not None:
            return n + 2
        return n

    def halo_bytes(self):
        x = self.ring.exchanger if self.ring is not None else (self.overlapped.exchanger if self.overlapped is not None else self.exchanger)
        return x.bytes_per_exchange if x is not None else 0

    def timed_run(self, steps, warmup, sampler=None):
        """W untimed steps, then exactly K steps between barrier + synchronize; CUDA events, max over ranks."""
        torch, dist = self.torch, self.dist
        for _ in range(warmup):
            self.step()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        single = self.world == 1
        kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)] if single else None

        def region():
            torch.cuda.synchronize()
            ev0.record()
            for s in range(steps):
                if single:
                    kev[s][0].record()
                    self.plan.execute()
                    kev[s][1].record()
                else:
                    self.step()
            ev1.record()
            torch.cuda.synchronize()
        if sampler is not None:
            with sampler:
                region()
        else:
            region()
        total_ms = ev0.elapsed_time(ev1)
        timeline = None
        stepper = self.ring if self.ring is not None else self.overlapped
        if stepper is not None:          # one more (untimed) step with event marks: where the exchange ends relative to the step
            tl = {}
            stepper.step()
            stepper.step(timeline=tl)
            torch.cuda.synchronize()
            timeline = stepper.timeline_ms(tl)
        if single:
            sweep_ms = sum(a.elapsed_time(b) for a, b in kev) / steps
        else:
            # the sweep kernel alone (ungated, no exchange), timed after the region
            if dist is not None:
                dist.barrier()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(steps):
                self.plan.execute()
            b.record()
            torch.cuda.synchronize()
            sweep_ms = a.elapsed_time(b) / steps
        per_rank = None
        t = torch.tensor([total_ms, sweep_ms], dtype=torch.float64, device=self.device)
        if dist is not None:
            keys = ("exchange_end", "interior_end", "strips_end", "end")
            mine = [total_ms / steps, sweep_ms] + [float((timeline or {}).get(k, 0.0)) for k in keys]
            every = [torch.zeros(6, dtype=torch.float64, device=self.device) for _ in range(self.world)]
            dist.all_gather(every, torch.tensor(mine, dtype=torch.float64, device=self.device))
            per_rank = [[round(float(x), 3) for x in e] for e in every]
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dist.barrier()
        total_ms, sweep_ms = float(t[0]), float(t[1])
        ms_per_step = total_ms / steps
        res = {"ms_per_step": ms_per_step, "sweep_ms": sweep_ms, "value": self.cells_job / (ms_per_step * 1e-3) / 1e9,
               "timeline": timeline, "per_rank": per_rank}
        for x in (self.ring, self.overlapped.exchanger if self.overlapped is not None else None, self.exchanger):
            if x is not None and hasattr(x, "check"):
                x.check()                                    # a halo wait that timed out invalidates the run
        return res

    def integrals(self):
        if self.ring is not None:
            return [float(x) for x in self.ring.total.cpu()]
        if self.overlapped is not None:
            return [float(x) for x in self.overlapped.total.cpu()]
        return self.grp.integrals()

    def parity_check(self):
        """N > 1: the step's integrals against the plain route -- NCCL send/recv halo exchange (2 rows), the ungated sweep of
        the whole slab, ncclAllReduce -- to 1e-12 of each integrand's scale (sum |term| V)."""
        torch, dist = self.torch, self.dist
        got = torch.tensor(self.integrals(), dtype=torch.float64, device=self.device)
        self.nccl_exchanger.exchange()
        self.plan.execute()
        want = self.plan.integrals[:got.numel()].clone()
        dist.all_reduce(want)
        scale = integrand_scales(self.ocb, torch, self.grid, self.m, dist)
        err = ((got - want).abs() / scale).cpu()
        spread = got.clone()
        dist.all_reduce(spread, op=dist.ReduceOp.MAX)
        same_bits = bool((spread == got).all())              # every rank holds the same integrals
        ok = bool((err <= 1e-12).all()) and bool(torch.isfinite(got).all()) and same_bits
        return {"parity_check": "ok" if ok else "FAILED", "against": "nccl halo exchange (2 rows) + ungated sweep + ncclAllReduce",
                "max_err_over_scale": float(err.max()), "tolerance": 1e-12, "identical_on_every_rank": same_bits,
                "reference_integrals": [float(x) for x in want.cpu()]}


def measured_traffic(size, fields):
    """DRAM bytes (read + write) of one launch of the dominant kernel, from the committed ncu capture of this workload
    (profiles/r2_traffic_bench1024_*.csv, `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum` over bench.py); None when
    no capture of this shape is committed."""
    import csv
    if tuple(size) != (1024, 1024, 1024):
        return None, None
    path = os.path.join(ROOT, "profiles", "r2_traffic_bench1024_%s.csv" % ("fields" if fields else "sums"))
    kern = "ocb_budget_kernel" if fields else "ocb_budget_sums_kernel"
    if not os.path.exists(path):
        return None, None
    per_id = {}
    with open(path, newline="") as f:
        rows = [r for r in csv.reader(f) if r]
    try:
        hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r and "Metric Value" in r)
    except StopIteration:
        return None, None
    h = rows[hdr]
    ci, ck, cm, cv, cu = h.index("ID"), h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit")
    mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    for r in rows[hdr + 1:]:
        if len(r) <= cv or kern not in r[ck] or not r[cm].startswith("dram__bytes_"):
            continue
        per_id[r[ci]] = per_id.get(r[ci], 0.0) + float(r[cv].replace(",", "")) * mult.get(r[cu], 1.0)
    if not per_id:
        return None, None
    return sum(per_id.values()) / len(per_id), os.path.relpath(path, ROOT)


def main():
    args = parse()
    if args.impl == "reference":
        from bench_reference import run_reference
        return run_reference(args)
    import torch
    import ocb_loader
    from bench_config import METRIC, UNIT, DEFAULT_SIZE, workload_config
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
        opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
        dist.init_process_group("nccl", device_id=device, pg_options=opts)
    size = tuple(args.size) if args.size else DEFAULT_SIZE
    if size[1] % world:
        raise SystemExit(f"Ny = {size[1]} is not divisible by {world} slabs")
    steps, warmup = args.steps, max(args.warmup, 3)

    # ---- headline: BASELINE configs[2] as written -- the `size` grid as ONE job, y slabs over the N GPUs (strong scaling) ----
    sampler = ClockSampler(local, enabled=(rank == 0))      # NVML set up here, outside the timed region; rank 0 reports its board
    run = Runner(ocb, torch, args, size, device, rank, world, dist)
    res = run.timed_run(steps, warmup, sampler)
    plan, grid = run.plan, run.grid
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = plan.algorithmic_bytes / (res["sweep_ms"] * 1e-3) / 1e9
    traffic, traffic_src = measured_traffic(size, args.fields) if world == 1 else (None, None)
    variant = {0: "generic", 1: "fused-budget", 2: "fused-budget + generic"}.get(plan.variant, str(plan.variant))
    out = {
        "metric": METRIC, "value": round(res["value"], 3), "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": round(res["ms_per_step"], 4), "higher_is_better": True, "scaling": "strong" if world > 1 else "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(size, world, args.fields),
        "kernel_variant": variant + (" (class sums)" if not args.fields else ""),
        "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                     "traffic": traffic, "traffic_source": traffic_src, "kernel_ms": round(res["sweep_ms"], 4),
                     "kernel": "ocb_budget_kernel" if args.fields else "ocb_budget_sums_kernel",
                     "algorithmic_bytes_per_cell": plan.algorithmic_bytes / run.cells_local, "cells_per_launch": run.cells_local,
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)"},
        "gpu_launches": run.launches_per_step() * steps,
        "clocks": sampler.summary(),
        "integrals": run.integrals(),
    }
    if world > 1:
        out["comm"] = {"step_minus_sweep_ms": round(res["ms_per_step"] - res["sweep_ms"], 4), "transport": run.transport,
                       "timeline_ms": res["timeline"],
                       "per_rank_ms": {"columns": ["loop_step", "sweep_alone", "exchange_end", "interior_end", "strips_end", "end"], "rows": res["per_rank"]},
                       "halo_bytes_per_rank_per_step": run.halo_bytes(),
                       "note": {"ring": "peer-memory ring, 3 launches per step: one kernel stores ONE edge row of the 5 inputs into each y "
                                        "neighbour's halo over NVLink and publishes the step; the class-sum sweep runs its two halo-reading "
                                        "tile rows last, gated in the kernel on the neighbours' flags; one block sums the partials and "
                                        "all-reduces the 7 integrals through the ranks' mapped sync blocks (no NCCL on the data path)",
                                "peer": "peer-memory push of 2 rows + stream-ordered wait, interior / strip sweeps, ncclAllReduce",
                                "nccl": "pack, grouped ncclSend/Recv ring, unpack, sweep, ncclAllReduce"}[run.transport]}
        out["comm"].update(run.parity_check())
    # ---- N > 1: the per-GPU-1024^3 (weak) run of round 1, as a sub-object -------------------------------------------
    if world > 1 and not args.no_weak:
        free_b, _ = torch.cuda.mem_get_info(device)
        need = 5 * (size[0] + 6) * (size[1] + 6) * (size[2] + 7) * 8 * 1.15
        del run, plan, grid
        import gc
        gc.collect()
        ok = torch.tensor([1.0 if free_b > need else 0.0], device=device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if float(ok) > 0:
            wjob = (size[0], size[1] * world, size[2])
            wrun = Runner(ocb, torch, args, wjob, device, rank, world, dist)
            wres = wrun.timed_run(steps, warmup)
            out["weak"] = {"scaling": "weak", "workload": f"{wjob[0]}x{wjob[1]}x{wjob[2]} Float64: {size[0]}x{size[1]}x{size[2]} per GPU",
                           "value": round(wres["value"], 3), "unit": UNIT, "ms_per_step": round(wres["ms_per_step"], 4),
                           "sweep_ms": round(wres["sweep_ms"], 4), "step_minus_sweep_ms": round(wres["ms_per_step"] - wres["sweep_ms"], 4),
                           "timeline_ms": wres["timeline"], "halo_bytes_per_rank_per_step": wrun.halo_bytes()}
            out["weak"].update(wrun.parity_check())
            run = wrun
        else:
            out["weak"] = {"skipped": "not enough free device memory for a 1024^3 slab per GPU next to the strong-scaling state"}
            run = Runner(ocb, torch, args, size, device, rank, world, dist)
    # ---- end to end with HOST inputs: every rank streams its own slab from pinned host memory -----------------------
    e2e = None
    if not args.no_e2e:
        if world > 1 and "weak" in out and "value" in out["weak"]:     # e2e is measured on the headline (strong) workload
            del run
            import gc
            gc.collect()
            run = Runner(ocb, torch, args, size, device, rank, world, dist)
        import psutil
        need = 5 * run.cells_local * 8 * 1.05 * world            # all ranks of this node pin their slab's five inputs
        ok = psutil.virtual_memory().available > 1.4 * need
        flag = torch.tensor([1.0 if ok else 0.0], device=device)
        if dist is not None:
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if float(flag) > 0:
            from bench_e2e import run_e2e
            e2e = run_e2e(ocb, torch, run.grid, run.m, args, dist=dist)
            t = torch.tensor([e2e["ms_per_step"]], dtype=torch.float64, device=device)
            if dist is not None:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e["ms_per_step"] = round(float(t[0]), 3)
            e2e["value"] = round(run.cells_job / (float(t[0]) * 1e-3) / 1e9, 4)
            e2e["h2d_bytes_per_step"] *= world
            e2e["d2h_bytes_per_step"] *= world
        else:
            e2e = {"value": None, "unit": UNIT, "h2d_bytes_per_step": None, "d2h_bytes_per_step": None,
                   "skipped": "not enough free host memory to pin every rank's inputs"}
    if e2e is not None:
        out["e2e"] = e2e
    # ---- N = 1: the other BASELINE configs, driver-timed as a `secondary` object --------------------------------------
    if world == 1 and not args.no_secondary:
        from bench_secondary import run_secondary
        out["secondary"] = run_secondary(ocb, torch, run, peak, steps=min(steps, 5))
    if rank == 0:
        if not args.no_cpu and world == 1:
            from bench_reference import cpu_baseline
            out["cpu_baseline"] = cpu_baseline(args)
        print(json.dumps(out))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
