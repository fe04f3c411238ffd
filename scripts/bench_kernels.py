#!/usr/bin/env python
"""Secondary measurements (not the driver's bench line): CUDA-event timings of the other kernels of the hot path at
the BASELINE.json configs -- config 2 (EPV + TKE budget, generic kernel), config 4 (stretched Float32 LES closures),
config 5 (Gaussian filter, FKE/SFKE graph, BPE sort).  Prints one JSON object per measurement."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ocb_loader  # noqa: E402

ocb = ocb_loader.load()
from bench import make_state  # noqa: E402


def timed(fn, warm=2, reps=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def report(name, ms, cells, bytes_per_cell, extra=None):
    out = {"name": name, "ms": round(ms, 3), "Gcells_per_s": round(cells / ms / 1e6, 3), "algorithmic_GBps": round(cells * bytes_per_cell / ms / 1e6, 1),
           "frac_of_6542": round(cells * bytes_per_cell / ms / 1e6 / 6542.4, 4), "bytes_per_cell": bytes_per_cell}
    out.update(extra or {})
    print(json.dumps(out), flush=True)


def main():
    dev = torch.device("cuda", 0)
    which = sys.argv[1:] or ["config2", "config4", "filter", "bpe"]
    if "config2" in which:
        n = 512
        grid, m = make_state(ocb, torch, (n, n, n), dev)
        m.coriolis = ocb.FPlane(1e-4)
        U, V, W = (ocb.average_xy(f) for f in (m.velocities.u, m.velocities.v, m.velocities.w))
        FD, KE, TKE = ocb.FlowDiagnostics, ocb.KineticEnergyEquation, ocb.TurbulentKineticEnergyEquation
        ops = [FD.ErtelPotentialVorticity(m), TKE.ShearProductionRate(m, U=U, V=V, W=W), KE.DissipationRate(m, U=U, V=V, W=W),
               KE.KineticEnergyPressureRedistribution(m)]
        grp = ocb.FusedDiagnostics(*[ocb.Integral(o) for o in ops])
        ms = timed(grp.plan.execute)
        report("config2: 512^3 F64 EPV + SP + eps(u') + PR, 4 integrals, one sweep", ms, n ** 3, grp.plan.algorithmic_bytes / n ** 3,
               {"variant": grp.plan.variant})
        del grp, m, U, V, W
        torch.cuda.empty_cache()
    for c4, dt4 in (("config4", torch.float32), ("config4f64", torch.float64)):
        if c4 not in which:
            continue
        import numpy as np
        Nx, Ny, Nz = 768, 768, 384
        k = np.arange(Nz + 1)
        zf = -np.tanh(2.0 * (1 - k / Nz)) / np.tanh(2.0)          # tanh-stretched faces, fine near the top
        grid = ocb.RectilinearGrid((Nx, Ny, Nz), x=(0, 1), y=(0, 1), z=zf, dtype=dt4, device=dev)
        nu_e = ocb.CenterField(grid)
        nu_e.interior.copy_(1e-3 * (1 + 0.5 * torch.rand(nu_e.interior.shape, device=dev)))
        nu_e.fill_halo_regions()
        for name, closure in (("SmagorinskyLilly-like nu_e field", ocb.Closure(nu_fields=(nu_e,))),
                              ("tuple (ScalarDiffusivity, AMD-like nu_e)", ocb.Closure.tuple_of(ocb.Closure(nu=1e-6), ocb.Closure(nu_fields=(nu_e,))))):
            m = ocb.NonhydrostaticModel(grid, closure=closure)
            for f in (m.velocities.u, m.velocities.v, m.velocities.w):
                f.interior.copy_(torch.randn(f.interior.shape, device=dev, dtype=dt4))
            m.fill_halo_regions()
            eps, smod = ocb.Field(ocb.KineticEnergyEquation.DissipationRate(m)), ocb.Field(ocb.FlowDiagnostics.StrainRateTensorModulus(m))
            grp = ocb.FusedDiagnostics(eps, smod)
            ms = timed(grp.plan.execute)
            report(f"{c4}: 768x768x384 {'F32' if dt4 == torch.float32 else 'F64'} stretched z, {name}: eps + |S| fields", ms, Nx * Ny * Nz, grp.plan.algorithmic_bytes / (Nx * Ny * Nz),
                   {"variant": grp.plan.variant})
            del grp, m, eps, smod
        torch.cuda.empty_cache()
    if "filter" in which or "bpe" in which or "filter_only" in which:
        Nx, Ny, Nz = 512, 512, 256
        grid = ocb.RectilinearGrid((Nx, Ny, Nz), extent=(1, 1, 0.5), device=dev)
        m = ocb.NonhydrostaticModel(grid, closure=ocb.Closure(nu=1e-6, kappa=1e-7))
        for f in (m.velocities.u, m.velocities.v, m.velocities.w):
            f.interior.copy_(torch.randn(f.interior.shape, device=dev, dtype=torch.float64))
        # buoyancy as SURVEY.md section 8d prescribes: b = N^2 z + 0.1 N^2 Lz xi
        zc = -0.5 + (torch.arange(Nz, device=dev, dtype=torch.float64) + 0.5) * (0.5 / Nz)
        m.tracers.b.interior.copy_(1e-5 * zc[:, None, None] + (0.1 * 1e-5 * 0.5) * torch.randn(m.tracers.b.interior.shape, device=dev, dtype=torch.float64))
        m.fill_halo_regions()
        cells = Nx * Ny * Nz
        if "filter_only" in which:        # the single staged filter alone (profiling runs)
            fu = ocb.Field(ocb.GaussianFilter(dims=(1, 2, 3), sigma=4.0 / 512, boundary="shrink")(m.velocities.u))
            report("config5: Gaussian filter dims=(1,2,3) sigma=4dx (17 taps) of one 512x512x256 F64 field, 3 staged passes", timed(fu.compute, warm=3, reps=10), cells, 16)
        if "filter" in which:
            sigma = 4.0 / 512
            filt = ocb.GaussianFilter(dims=(1, 2, 3), sigma=sigma, boundary="shrink")
            fu = ocb.Field(filt(m.velocities.u))
            ms = timed(fu.compute)
            report("config5: Gaussian filter dims=(1,2,3) sigma=4dx (17 taps) of one 512x512x256 F64 field, 3 staged passes", ms, cells, 16,
                   {"note": "strict algorithmic bytes 2*8 B/cell; 3 passes move 6*8 B/cell"})
            FKE, SFKE = ocb.FilteredKineticEnergyEquation, ocb.SubFilterKineticEnergyEquation
            for name, ctor in (("FilteredKineticEnergy", FKE.FilteredKineticEnergy), ("KineticEnergyCrossScaleFlux", FKE.KineticEnergyCrossScaleFlux),
                               ("FilteredKineticEnergyDissipationRate", FKE.FilteredKineticEnergyDissipationRate),
                               ("SubFilterKineticEnergy", SFKE.SubFilterKineticEnergy), ("SubFilterKineticEnergyDissipationRate", SFKE.SubFilterKineticEnergyDissipationRate)):
                f = ocb.Field(ctor(m, filt))
                t = [0]
                def run(f=f, t=t):
                    t[0] += 1
                    f.compute(time=float(t[0]))
                ms = timed(run, warm=1, reps=2)
                report(f"config5: {name} (whole graph: filters + contraction)", ms, cells, 0)
                del f
                torch.cuda.empty_cache()
        if "bpe" in which:
            BPE = ocb.BackgroundPotentialEnergyEquation
            for method in (BPE.ThreeDimensionalSort(), BPE.HeavisideIntegral()):
                z = BPE.reference_height(m.tracers.b, method=method)
                ms = timed(z.compute, warm=1, reps=3)
                report(f"config5: reference_height {method!r} on 512x512x256 F64 (67.1 M keys)", ms, cells, 248,
                       {"Mkeys_per_s": round(cells / ms / 1e3, 1), "note": "248 B/key onesweep model (SURVEY 8d)"})
            keys = torch.randn(cells, device=dev, dtype=torch.float64)
            out = torch.empty_like(keys)
            perm = torch.empty(cells, dtype=torch.int32, device=dev)
            L = ocb._lib.lib()
            ms = timed(lambda: ocb._lib.check(L.ocb_sort_pairs(1, keys.data_ptr(), out.data_ptr(), perm.data_ptr(), cells, torch.cuda.current_stream().cuda_stream)))
            report("radix sort alone: 67.1 M Float64 keys + 32-bit payload", ms, cells, 8 * 2 * 12 + 8, {"Mkeys_per_s": round(cells / ms / 1e3, 1)})


if __name__ == "__main__":
    main()
