#!/usr/bin/env python
"""KE + tracer-variance budget on a stretched-z grid (512^3 Float64, constant coefficients): the per-plane-metric variant of the
budget kernel against the generic kernel (OCB_DISABLE_FUSED_STRETCHED=1), 7 integrals and 7 integrals + 7 fields."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ocb_loader  # noqa: E402

ocb = ocb_loader.load()
from bench import budget_ops  # noqa: E402


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


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    dev = torch.device("cuda", 0)
    k = np.arange(n + 1)
    zf = -np.tanh(2.0 * (1 - k / n)) / np.tanh(2.0)
    grid = ocb.RectilinearGrid((n, n, n), x=(0, 1), y=(0, 1), z=zf, device=dev)
    m = ocb.NonhydrostaticModel(grid, closure=ocb.Closure(nu=1e-6, kappa=1e-7))
    for f in (m.velocities.u, m.velocities.v, m.velocities.w, m.tracers.b, m.pressures.pNHS):
        f.interior.copy_(torch.randn(f.interior.shape, device=dev, dtype=torch.float64))
    m.fill_halo_regions()
    for fields in (False, True):
        for generic in (False, True):
            if generic:
                os.environ["OCB_DISABLE_FUSED_STRETCHED"] = "1"
            try:
                ops = budget_ops(ocb, m)
                grp = ocb.FusedDiagnostics(*([ocb.Integral(o) for o in ops] + ([ocb.Field(o) for o in ops] if fields else [])))
            finally:
                os.environ.pop("OCB_DISABLE_FUSED_STRETCHED", None)
            ms = timed(grp.plan.execute, reps=3 if generic else 5)
            cells = n ** 3
            print(json.dumps({"name": f"stretched-z budget {n}^3, 7 integrals" + (" + 7 fields" if fields else ""), "kernel": "generic" if generic else "budget kernel (per-plane metrics)",
                              "variant": grp.plan.variant, "ms": round(ms, 3), "Gcells_per_s": round(cells / ms / 1e6, 2),
                              "frac_of_6542": round(grp.plan.algorithmic_bytes / ms / 1e6 / 6542.4, 4), "integrals": [float("%.6e" % v) for v in grp.integrals()[:3]]}), flush=True)
            del grp, ops
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
