"""bench.py's `secondary` object (N = 1): the BASELINE.json configs other than the headline, each timed with CUDA events
(warm-up 3, then `steps` executions) against its own roofline -- algorithmic bytes per cell over the
measured HBM peak.  Synthetic inputs as in the headline run; every entry names the kernels that ran (`variant`)."""
from __future__ import annotations

import gc

import numpy as np


def _timed(torch, fn, warm=3, reps=5):
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


def _entry(ms, cells, bytes_per_cell, peak, **extra):
    gbs = cells * bytes_per_cell / ms / 1e6
    out = {"ms": round(ms, 4), "Gcells_per_s": round(cells / ms / 1e6, 3), "algorithmic_bytes_per_cell": bytes_per_cell,
           "achieved_GBps": round(gbs, 1), "frac": round(gbs / peak, 4)}
    out.update(extra)
    return out


VARIANTS = {0: "generic", 1: "budget kernel", 2: "budget + generic", 3: "velocity-gradient kernel", 4: "velocity-gradient + generic",
            5: "budget + velocity-gradient", 6: "budget + velocity-gradient + generic"}


def run_secondary(ocb, torch, run, peak, steps=5):
    from bench import budget_ops, make_state
    dev = run.device
    out = {"peak_GBps": peak, "timing": f"CUDA events, 3 warm-up + {steps} timed executions each"}
    # ---- the headline workload plus the TKE budget about U(z), V(z), W(z): north_star's "TKE + KE + tracer-variance" plan -------
    try:
        m = run.m
        U, V, W = (ocb.average_xy(f) for f in (m.velocities.u, m.velocities.v, m.velocities.w))
        TKE, KE = ocb.TurbulentKineticEnergyEquation, ocb.KineticEnergyEquation
        ops = budget_ops(ocb, m) + [TKE.ShearProductionRate(m, U=U, V=V, W=W), KE.DissipationRate(m, U=U, V=V, W=W),
                                    TKE.TurbulentKineticEnergy(m, U=U, V=V, W=W)]
        grp = ocb.FusedDiagnostics(*[ocb.Integral(o) for o in ops])
        cells = run.cells_local
        ms = _timed(torch, grp.plan.execute, reps=steps)
        out["config3A_tke_ke_tracer_variance"] = _entry(
            ms, cells, grp.plan.algorithmic_bytes / cells, peak, variant=VARIANTS.get(grp.plan.variant), launches=grp.plan.n_launches,
            workload="1024^3 F64: KE budget {K, ADV, PR, u_i b_i, eps} + tracer variance {chi, 2c div F} + TKE budget about "
                     "U(z), V(z), W(z) {shear production, eps(u - U), TKE}: 10 volume integrals, one sweep", integrals=grp.integrals())
        del grp, ops, U, V, W
    except Exception as e:  # pragma: no cover
        out["config3A_tke_ke_tracer_variance"] = {"error": repr(e)[:300]}
    # ---- config 3B: the headline workload with the 7 terms also written as fields (96 B/cell) ---------------------------
    try:
        ops = budget_ops(ocb, run.m)
        grp = ocb.FusedDiagnostics(*([ocb.Integral(o) for o in ops] + [ocb.Field(o) for o in ops]))
        cells = run.cells_local
        ms = _timed(torch, grp.plan.execute, reps=steps)
        out["config3B"] = _entry(ms, cells, grp.plan.algorithmic_bytes / cells, peak, variant=VARIANTS.get(grp.plan.variant),
                                 workload="1024^3 F64: 7 integrals + 7 output fields, one sweep")
        del grp, ops
    except Exception as e:  # pragma: no cover
        out["config3B"] = {"error": repr(e)[:300]}
    # release the 1024^3 state
    run.grp = run.plan = run.m = run.fields = run.grid = None
    gc.collect()
    torch.cuda.empty_cache()
    # ---- config 2: 512^3 EPV + TKE budget about horizontal-mean profiles, 4 integrals, one sweep -------------------------
    try:
        n = 512
        grid, m = make_state(ocb, torch, (n, n, n), dev)
        m.coriolis = ocb.FPlane(1e-4)
        U, V, W = (ocb.average_xy(f) for f in (m.velocities.u, m.velocities.v, m.velocities.w))
        FD, KE, TKE = ocb.FlowDiagnostics, ocb.KineticEnergyEquation, ocb.TurbulentKineticEnergyEquation
        ops = [FD.ErtelPotentialVorticity(m), TKE.ShearProductionRate(m, U=U, V=V, W=W), KE.DissipationRate(m, U=U, V=V, W=W),
               KE.KineticEnergyPressureRedistribution(m)]
        grp = ocb.FusedDiagnostics(*[ocb.Integral(o) for o in ops])
        ms = _timed(torch, grp.plan.execute, reps=steps)
        out["config2"] = _entry(ms, n ** 3, grp.plan.algorithmic_bytes / n ** 3, peak, variant=VARIANTS.get(grp.plan.variant),
                                workload="512^3 F64: EPV + SP + eps(u') + PR about U(z), V(z), W(z), 4 integrals, one sweep")
        del grp, ops, m, U, V, W, grid
    except Exception as e:  # pragma: no cover
        out["config2"] = {"error": repr(e)[:300]}
    gc.collect()
    torch.cuda.empty_cache()
    # ---- the KE + tracer-variance budget on a stretched z axis (the grid of config 4 in Float64, constant coefficients) ----
    try:
        n = 512
        k = np.arange(n + 1)
        zf = -np.tanh(2.0 * (1 - k / n)) / np.tanh(2.0)
        grid = ocb.RectilinearGrid((n, n, n), x=(0, 1), y=(0, 1), z=zf, device=dev)
        m = ocb.NonhydrostaticModel(grid, closure=ocb.Closure(nu=1e-6, kappa=1e-7))
        for f in (m.velocities.u, m.velocities.v, m.velocities.w, m.tracers.b, m.pressures.pNHS):
            f.interior.copy_(torch.randn(f.interior.shape, device=dev, dtype=torch.float64))
        m.fill_halo_regions()
        ops = budget_ops(ocb, m)
        for key, members, what in (("stretched_budget_integrals", [ocb.Integral(o) for o in ops], "7 integrals"),
                                   ("stretched_budget_fields", [ocb.Integral(o) for o in ops] + [ocb.Field(o) for o in ops], "7 integrals + 7 fields")):
            grp = ocb.FusedDiagnostics(*members)
            ms = _timed(torch, grp.plan.execute, reps=steps)
            out[key] = _entry(ms, n ** 3, grp.plan.algorithmic_bytes / n ** 3, peak, variant=VARIANTS.get(grp.plan.variant),
                              workload=f"512^3 F64, stretched z (tanh faces), constant nu / kappa: KE + tracer-variance budget, {what}, one sweep "
                                       "(per-plane-metric variants of the class-sum / budget kernels)")
            del grp
        # the same grid with a SmagorinskyLilly-like closure (nu_e field, kappa_e = nu_e / Pr): the LES budget integrals in one launch
        nu_e = ocb.CenterField(grid)
        nu_e.interior.copy_(1e-3 * (1 + 0.5 * torch.rand(nu_e.interior.shape, device=dev, dtype=torch.float64)))
        nu_e.fill_halo_regions()
        mles = ocb.NonhydrostaticModel(grid, closure=ocb.Closure(nu_fields=(nu_e,), kappa_fields=(nu_e,), kappa_scale=(1.0,), name="SmagorinskyLilly"))
        for f, g_ in zip((mles.velocities.u, mles.velocities.v, mles.velocities.w, mles.tracers.b, mles.pressures.pNHS),
                         (m.velocities.u, m.velocities.v, m.velocities.w, m.tracers.b, m.pressures.pNHS)):
            f.data.copy_(g_.data)
        grp = ocb.FusedDiagnostics(*[ocb.Integral(o) for o in budget_ops(ocb, mles)])
        ms = _timed(torch, grp.plan.execute, reps=steps)
        out["stretched_les_budget_integrals"] = _entry(
            ms, n ** 3, grp.plan.algorithmic_bytes / n ** 3, peak, variant=VARIANTS.get(grp.plan.variant), launches=grp.plan.n_launches,
            workload="512^3 F64, stretched z, eddy-viscosity closure (nu_e field, kappa_e = nu_e / Pr): the 7 budget integrals, one sweep "
                     "(class sums with per-item viscosities / diffusivities)")
        del grp, mles, nu_e
        del ops, m, grid
    except Exception as e:  # pragma: no cover
        out["stretched_budget"] = {"error": repr(e)[:300]}
    gc.collect()
    torch.cuda.empty_cache()
    # ---- config 4: 768x768x384 Float32, stretched z, eddy-viscosity closures: eps + |S| fields ---------------------------
    try:
        Nx, Ny, Nz = 768, 768, 384
        k = np.arange(Nz + 1)
        zf = -np.tanh(2.0 * (1 - k / Nz)) / np.tanh(2.0)
        grid = ocb.RectilinearGrid((Nx, Ny, Nz), x=(0, 1), y=(0, 1), z=zf, dtype=torch.float32, device=dev)
        nu_e = ocb.CenterField(grid)
        nu_e.interior.copy_(1e-3 * (1 + 0.5 * torch.rand(nu_e.interior.shape, device=dev)))
        nu_e.fill_halo_regions()
        for key, name, closure in (("config4_smagorinsky", "SmagorinskyLilly-like nu_e field", ocb.Closure(nu_fields=(nu_e,))),
                                   ("config4_tuple", "tuple (ScalarDiffusivity, AMD-like nu_e field)",
                                    ocb.Closure.tuple_of(ocb.Closure(nu=1e-6), ocb.Closure(nu_fields=(nu_e,))))):
            m = ocb.NonhydrostaticModel(grid, closure=closure)
            for f in (m.velocities.u, m.velocities.v, m.velocities.w):
                f.interior.copy_(torch.randn(f.interior.shape, device=dev, dtype=torch.float32))
            m.fill_halo_regions()
            eps, smod = ocb.Field(ocb.KineticEnergyEquation.DissipationRate(m)), ocb.Field(ocb.FlowDiagnostics.StrainRateTensorModulus(m))
            grp = ocb.FusedDiagnostics(eps, smod)
            ms = _timed(torch, grp.plan.execute, reps=steps)
            out[key] = _entry(ms, Nx * Ny * Nz, grp.plan.algorithmic_bytes / (Nx * Ny * Nz), peak, variant=VARIANTS.get(grp.plan.variant),
                              workload=f"768x768x384 F32 stretched z, {name}: eps + |S| fields")
            del grp, eps, smod
            if key == "config4_smagorinsky":
                # the same Float32 LES state: the seven KE + tracer-variance budget integrals in one launch (Float32-input class sums)
                for f in (m.tracers.b, m.pressures.pNHS):
                    f.interior.copy_(torch.randn(f.interior.shape, device=dev, dtype=torch.float32))
                m.fill_halo_regions()
                grp = ocb.FusedDiagnostics(*[ocb.Integral(o) for o in budget_ops(ocb, m)])
                ms = _timed(torch, grp.plan.execute, reps=steps)
                out["config4_les_budget_integrals"] = _entry(
                    ms, Nx * Ny * Nz, grp.plan.algorithmic_bytes / (Nx * Ny * Nz), peak, variant=VARIANTS.get(grp.plan.variant), launches=grp.plan.n_launches,
                    workload="768x768x384 F32 stretched z, SmagorinskyLilly-like nu_e field: K, ADV, PR, w b, eps, chi, 2c div(kappa grad c) integrals, one sweep")
                del grp
            del m
        del nu_e, grid
    except Exception as e:  # pragma: no cover
        out["config4"] = {"error": repr(e)[:300]}
    gc.collect()
    torch.cuda.empty_cache()
    # ---- config 5: 512x512x256 F64 Gaussian filter, FKE / SFKE graphs, sort-based reference height -------------------------
    try:
        Nx, Ny, Nz = 512, 512, 256
        cells = Nx * Ny * Nz
        grid = ocb.RectilinearGrid((Nx, Ny, Nz), extent=(1, 1, 0.5), device=dev)
        m = ocb.NonhydrostaticModel(grid, closure=ocb.Closure(nu=1e-6, kappa=1e-7))
        for f in (m.velocities.u, m.velocities.v, m.velocities.w):
            f.interior.copy_(torch.randn(f.interior.shape, device=dev, dtype=torch.float64))
        # buoyancy as SURVEY.md section 8d prescribes: b = N^2 z + 0.1 N^2 Lz xi (a stratified field with noise, not white noise)
        zc = -0.5 + (torch.arange(Nz, device=dev, dtype=torch.float64) + 0.5) * (0.5 / Nz)
        m.tracers.b.interior.copy_(1e-5 * zc[:, None, None] + (0.1 * 1e-5 * 0.5) * torch.randn(m.tracers.b.interior.shape, device=dev, dtype=torch.float64))
        m.fill_halo_regions()
        filt = ocb.GaussianFilter(dims=(1, 2, 3), sigma=4.0 / 512, boundary="shrink")
        fu = ocb.Field(filt(m.velocities.u))
        ms = _timed(torch, fu.compute, reps=steps)
        out["config5_filter"] = _entry(ms, cells, 16, peak, workload="Gaussian filter dims=(1,2,3), sigma = 4 dx (17 taps, shrink in z) of one "
                                       "512x512x256 F64 field; strict 2 x 8 B/cell", passes=getattr(fu, "n_kernel_passes", None))
        FKE, SFKE = ocb.FilteredKineticEnergyEquation, ocb.SubFilterKineticEnergyEquation
        graphs = {}
        for name, ctor in (("FilteredKineticEnergy", FKE.FilteredKineticEnergy), ("KineticEnergyCrossScaleFlux", FKE.KineticEnergyCrossScaleFlux),
                           ("FilteredKineticEnergyDissipationRate", FKE.FilteredKineticEnergyDissipationRate),
                           ("SubFilterKineticEnergy", SFKE.SubFilterKineticEnergy),
                           ("SubFilterKineticEnergyDissipationRate", SFKE.SubFilterKineticEnergyDissipationRate)):
            f = ocb.Field(ctor(m, filt))
            t = [0]

            def go(f=f, t=t):
                t[0] += 1
                f.compute(time=float(t[0]))
            graphs[name] = round(_timed(torch, go, warm=1, reps=2), 3)
            del f
            torch.cuda.empty_cache()
        out["config5_fke_sfke_graphs_ms"] = graphs
        BPE = ocb.BackgroundPotentialEnergyEquation
        for key, method in (("config5_bpe_sort", BPE.ThreeDimensionalSort()), ("config5_bpe_heaviside", BPE.HeavisideIntegral())):
            z = BPE.reference_height(m.tracers.b, method=method)
            ms = _timed(torch, z.compute, warm=2, reps=steps)
            out[key] = _entry(ms, cells, 248, peak, Mkeys_per_s=round(cells / ms / 1e3, 1),
                              workload=f"reference_height {method!r} on 512x512x256 F64 (67.1 M keys); 248 B/key onesweep model")
            del z
        keys = torch.randn(cells, device=dev, dtype=torch.float64)
        ko = torch.empty_like(keys)
        perm = torch.empty(cells, dtype=torch.int32, device=dev)
        L = ocb._lib.lib()
        ms = _timed(torch, lambda: ocb._lib.check(L.ocb_sort_pairs(1, keys.data_ptr(), ko.data_ptr(), perm.data_ptr(), cells,
                                                                   torch.cuda.current_stream().cuda_stream)), warm=2, reps=steps)
        out["config5_sort_alone"] = _entry(ms, cells, 200, peak, Mkeys_per_s=round(cells / ms / 1e3, 1),
                                           workload="radix sort of 67.1 M Float64 keys + 32-bit payload; 8 x 2 x 12 + 8 B/key")
    except Exception as e:  # pragma: no cover
        out["config5"] = {"error": repr(e)[:300]}
    gc.collect()
    torch.cuda.empty_cache()
    return out
