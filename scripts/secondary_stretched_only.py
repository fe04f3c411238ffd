import sys, os, json, torch
sys.path.insert(0, os.getcwd())
import ocb_loader
ocb = ocb_loader.load()
import bench_secondary as B
class R: pass
r = R(); r.device = torch.device("cuda", 0)
# minimal runner: only the stretched block is of interest; emulate by calling run_secondary with a tiny headline state
from bench import make_state
grid, m = make_state(ocb, torch, (128, 128, 128), r.device)
r.m = m; r.cells_local = 128 ** 3; r.grp = r.plan = r.fields = r.grid = None
out = B.run_secondary(ocb, torch, r, 6542.4, steps=3)
for k in ("stretched_budget_integrals", "stretched_budget_fields", "stretched_les_budget_integrals", "stretched_budget"):
    if k in out: print(k, json.dumps(out[k])[:300])
