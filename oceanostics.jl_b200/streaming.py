"""Host-resident fields streamed through the GPU: the end-to-end form of the fused budget sweep.

When the model state lives in HOST memory (an Oceananigans CPU model, or output read back from disk), every
evaluation has to bring the fields to the device first.  `HostStreamedBudget` does that as a pipeline: the domain is
cut into z slabs; a copy stream moves slab c+1 of every input (pinned host -> device) while the compute stream runs
the fused sweep on slab c; the per-slab volume integrals are summed on the device and read back once.  The numbers
are the same fused-kernel results (`Integral(op, indices=...)` per slab), only the data path differs.
"""
from __future__ import annotations

import torch

from .operations import FusedDiagnostics, Integral


class HostStreamedBudget:
    def __init__(self, grid, device_fields, ops, n_slabs=16, pin=True):
        """device_fields: the device Fields the operations `ops` read (their parents are the copy destinations);
        ops: KernelFunctionOperations evaluated as volume integrals."""
        if grid.device.type != "cuda":
            raise RuntimeError("HostStreamedBudget needs a CUDA grid")
        self.grid, self.fields, self.ops = grid, list(device_fields), list(ops)
        Nz, H = grid.N[2], grid.H[2]
        n_slabs = max(1, min(n_slabs, Nz))
        edges = [round(c * Nz / n_slabs) for c in range(n_slabs + 1)]
        self.slabs = [(edges[c] + 1, edges[c + 1]) for c in range(n_slabs) if edges[c + 1] > edges[c]]   # 1-based k ranges
        # host mirrors of the parent arrays (pinned, so the copies are asynchronous DMA)
        self.host = [torch.empty(f.data.shape, dtype=f.data.dtype, pin_memory=pin) for f in self.fields]
        self.groups = [FusedDiagnostics(*[Integral(o, indices=(None, None, (k0, k1))) for o in self.ops]) for (k0, k1) in self.slabs]
        self.copy_stream = torch.cuda.Stream(device=grid.device)
        self.events = [torch.cuda.Event() for _ in self.slabs]
        self.total = torch.zeros(len(self.ops), dtype=torch.float64, device=grid.device)
        self.result_host = torch.empty(len(self.ops), dtype=torch.float64, pin_memory=pin)
        # parent-plane range each slab's sweep reads beyond the previous slab: planes k0-1 .. k1+2 (1-based k -> parent k+H-1)
        self.plane_ranges = []
        done = 0
        top = max(f.data.shape[0] for f in self.fields)
        for (k0, k1) in self.slabs:
            hi = min(k1 + 2 + H - 1, top - 1)
            if (k0, k1) == self.slabs[-1]:
                hi = top - 1
            self.plane_ranges.append((done, hi + 1))
            done = hi + 1
        self.h2d_bytes = sum(h.numel() * h.element_size() for h in self.host)
        self.d2h_bytes = self.result_host.numel() * 8

    def load_host_from_device(self):
        """Initialise the host mirrors with the current device state (set-up, not part of a step)."""
        for h, f in zip(self.host, self.fields):
            h.copy_(f.data)
        torch.cuda.synchronize()

    def step(self):
        """One end-to-end evaluation: H2D of every input, fused sweeps per slab, D2H of the integrals."""
        cur = torch.cuda.current_stream(self.grid.device)
        self.copy_stream.wait_stream(cur)          # the previous step's sweeps are done with the destination planes
        with torch.cuda.stream(self.copy_stream):
            for c, (p0, p1) in enumerate(self.plane_ranges):
                for h, f in zip(self.host, self.fields):
                    q1 = min(p1, f.data.shape[0])
                    if q1 > p0:
                        f.data[p0:q1].copy_(h[p0:q1], non_blocking=True)
                self.events[c].record(self.copy_stream)
        self.total.zero_()
        for c, grp in enumerate(self.groups):
            cur.wait_event(self.events[c])
            grp.plan.execute()
            self.total.add_(grp.plan.integrals[:len(self.ops)])
        self.result_host.copy_(self.total, non_blocking=True)
        return self.result_host

    def integrals(self):
        torch.cuda.current_stream(self.grid.device).synchronize()
        return [float(x) for x in self.result_host]
