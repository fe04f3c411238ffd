#!/usr/bin/env python
"""The ten-integral TKE + KE + tracer-variance plan at 1024^3 (bench secondary's config3A entry) alone, for shape tuning runs."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ocb_loader
ocb = ocb_loader.load()
from bench import make_state, budget_ops
dev = torch.device("cuda", 0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
grid, m = make_state(ocb, torch, (n, n, n), dev)
U, V, W = (ocb.average_xy(f) for f in (m.velocities.u, m.velocities.v, m.velocities.w))
TKE, KE = ocb.TurbulentKineticEnergyEquation, ocb.KineticEnergyEquation
ops = budget_ops(ocb, m) + [TKE.ShearProductionRate(m, U=U, V=V, W=W), KE.DissipationRate(m, U=U, V=V, W=W), TKE.TurbulentKineticEnergy(m, U=U, V=V, W=W)]
grp = ocb.FusedDiagnostics(*[ocb.Integral(o) for o in ops])
for _ in range(3): grp.plan.execute()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): grp.plan.execute()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(json.dumps({"ms": round(ms, 3), "frac": round(n ** 3 * 40 / ms / 1e6 / 6542.4, 4), "variant": grp.plan.variant}))
