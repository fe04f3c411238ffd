#!/usr/bin/env python
"""bench_secondary's config-4 block alone (profiling / timing runs)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ocb_loader  # noqa: E402

ocb = ocb_loader.load()
import bench_secondary as B  # noqa: E402
from bench import make_state  # noqa: E402


class R:
    pass


r = R()
r.device = torch.device("cuda", 0)
grid, m = make_state(ocb, torch, (128, 128, 128), r.device)
r.m = m; r.cells_local = 128 ** 3; r.grp = r.plan = r.fields = r.grid = None
out = B.run_secondary(ocb, torch, r, 6542.4, steps=3)
for k, v in out.items():
    if k.startswith("config4") or k.startswith("stretched"):
        print(k, json.dumps(v)[:260])
