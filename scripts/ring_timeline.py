#!/usr/bin/env python
"""Per-step timeline of the ring step under torchrun: CUDA events at the start and end of every step and at the end of its
push, plus the host's enqueue times -- tells GPU time inside a step from gaps between steps (host-bound enqueue)."""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ocb_loader  # noqa: E402

ocb = ocb_loader.load()
from bench import make_state, budget_ops, slab_inputs  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from oceanostics_b200 import distributed as D
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    grid, m = make_state(ocb, torch, (n, n, n), dev, rank, world)
    flds = slab_inputs(m)
    D.SlabHaloExchanger(grid, flds, depth=2, rank=rank, world=world).exchange()
    for f in flds:
        f.fill_halo_regions_xz()
    rb = D.RingSlabBudget(grid, flds, lambda: budget_ops(ocb, m), rank=rank, world=world)
    for _ in range(5):
        rb.step()
    torch.cuda.synchronize(); dist.barrier()
    tls, host = [], []
    t0 = time.perf_counter()
    for _ in range(steps):
        tl = {}
        host.append(time.perf_counter() - t0)
        rb.step(timeline=tl)
        tls.append(tl)
    host.append(time.perf_counter() - t0)
    torch.cuda.synchronize()
    first = tls[0]["start"]
    rows = [[round(first.elapsed_time(tl[k]), 3) for k in ("start", "exchange_end", "end")] for tl in tls]
    # the ungated sweep and the gated sweep + reduction without pushes in flight
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        rb.plan.execute()
    b.record()
    torch.cuda.synchronize()
    alone = a.elapsed_time(b) / steps
    out = {"rank": rank, "sweep_alone_ms": round(alone, 3), "step_ms": [round(r[2] - r[0], 3) for r in rows],
           "gap_ms": [round(rows[i + 1][0] - rows[i][2], 3) for i in range(len(rows) - 1)],
           "push_end_ms": [round(r[1] - r[0], 3) for r in rows], "host_enqueue_ms": [round((host[i + 1] - host[i]) * 1e3, 3) for i in range(steps)]}
    for r in range(world):
        if r == rank:
            print(json.dumps(out), flush=True)
        dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
