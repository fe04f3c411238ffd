"""Times the three y windows of the overlapped slab step (edge strips + interior rows) on one GPU, for the class-sum kernel
and for the cell-exact integral mode (OCB_FUSED_CELL_INTEGRALS=1), at the bench workload."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import ocb_loader

ocb = ocb_loader.load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
grid, m = bench.make_state(ocb, torch, (n, n, n), torch.device("cuda", 0))
ops = bench.budget_ops(ocb, m)


def timed(plan, reps=10):
    for _ in range(3):
        plan.execute()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        plan.execute()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


for mode in ("sums", "cells"):
    if mode == "cells":
        os.environ["OCB_FUSED_CELL_INTEGRALS"] = "1"
    tot = None
    for w in ((1, 2), (3, n - 2), (n - 1, n), (1, n)):
        grp = ocb.FusedDiagnostics(*[ocb.Integral(o, indices=(None, w, None)) for o in ops])
        ms = timed(grp.plan)
        vals = grp.compute().integrals()
        print(mode, w, "variant", grp.plan.variant, f"{ms:.4f} ms", [f"{v:.15e}" for v in vals[:2]], flush=True)
    os.environ.pop("OCB_FUSED_CELL_INTEGRALS", None)
