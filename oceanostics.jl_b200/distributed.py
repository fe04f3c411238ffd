"""Slab decomposition across the GPUs of one node: one rank per GPU, the domain cut into y slabs (the partition
Oceananigans' Distributed architecture supports; the contiguous x axis stays whole), periodic ring in y.

Per sweep the only data-path communication is (1) the exchange of `depth` y-planes of every input field with the two
ring neighbours -- packed by the library's pack kernel into contiguous buffers, moved with grouped NCCL send/recv over
NVLink, unpacked into the neighbour's halo -- and (2) one all-reduce of the vector of volume integrals.  The sweep
itself needs no collective.  The reference has no distributed layer of its own.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.distributed as dist

from . import _lib
from .fields import current_stream_ptr


def slab_range(n_global, rank, world):
    """1-based inclusive global y range owned by `rank` (even split; the remainder goes to the low ranks)."""
    base, rem = divmod(n_global, world)
    lo = rank * base + min(rank, rem)
    return lo + 1, lo + base + (1 if rank < rem else 0)


def _pack_cuda(field, side, h, buf):
    L = _lib.lib()
    d = field.descriptor()
    dt = _lib.OCB_F64 if field.data.dtype == torch.float64 else _lib.OCB_F32
    _lib.check(L.ocb_halo_pack_y(dt, C.byref(d), side, h, buf.data_ptr(), current_stream_ptr(field.data.device)))


def _unpack_cuda(field, side, h, buf):
    L = _lib.lib()
    d = field.descriptor()
    dt = _lib.OCB_F64 if field.data.dtype == torch.float64 else _lib.OCB_F32
    _lib.check(L.ocb_halo_unpack_y(dt, C.byref(d), side, h, buf.data_ptr(), current_stream_ptr(field.data.device)))


class SlabHaloExchanger:
    """Exchange `depth` y-planes of each field with the ring neighbours.

    pack / unpack default to the library's CUDA kernels; the CPU (gloo) tests of the exchange schedule inject
    torch-slicing packers because the library has no CPU path."""

    def __init__(self, grid, fields, depth=2, rank=None, world=None, group=None, pack=None, unpack=None):
        self.grid, self.fields, self.h, self.group = grid, list(fields), int(depth), group
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world = dist.get_world_size(group) if world is None else world
        self.pack = pack or _pack_cuda
        self.unpack = unpack or _unpack_cuda
        if self.h < 1 or self.h > grid.H[1]:
            raise ValueError(f"exchange depth {depth} must be within 1..halo ({grid.H[1]})")
        self.lo_rank = (self.rank - 1) % self.world
        self.hi_rank = (self.rank + 1) % self.world

        def buf(f):
            return torch.empty((f.data.shape[0], self.h, f.data.shape[2]), dtype=f.data.dtype, device=f.data.device)

        self.send_lo = [buf(f) for f in self.fields]
        self.send_hi = [buf(f) for f in self.fields]
        self.recv_lo = [buf(f) for f in self.fields]
        self.recv_hi = [buf(f) for f in self.fields]
        self.bytes_per_exchange = sum(2 * b.numel() * b.element_size() for b in self.send_lo)

    def exchange(self):
        h = self.h
        for f, slo, shi in zip(self.fields, self.send_lo, self.send_hi):
            self.pack(f, 0, h, slo)       # my low-edge interior planes -> the low neighbour's high halo
            self.pack(f, 1, h, shi)       # my high-edge interior planes -> the high neighbour's low halo
        if self.world == 1:
            for slo, shi, rlo, rhi in zip(self.send_lo, self.send_hi, self.recv_lo, self.recv_hi):
                rhi.copy_(slo)
                rlo.copy_(shi)
        else:
            ops = []
            for slo, shi, rlo, rhi in zip(self.send_lo, self.send_hi, self.recv_lo, self.recv_hi):
                ops.append(dist.P2POp(dist.isend, slo, self.lo_rank, self.group))
                ops.append(dist.P2POp(dist.isend, shi, self.hi_rank, self.group))
                ops.append(dist.P2POp(dist.irecv, rhi, self.hi_rank, self.group))
                ops.append(dist.P2POp(dist.irecv, rlo, self.lo_rank, self.group))
            for req in dist.batch_isend_irecv(ops):   # one grouped NCCL launch for all fields and both directions
                req.wait()
        for f, rlo, rhi in zip(self.fields, self.recv_lo, self.recv_hi):
            self.unpack(f, 0, h, rlo)
            self.unpack(f, 1, h, rhi)


def allreduce_integrals(integrals: torch.Tensor, group=None):
    """Global `Integral`s: sum of the per-slab partial integrals (ncclAllReduce of K doubles)."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(integrals, op=dist.ReduceOp.SUM, group=group)
    return integrals


class OverlappedSlabBudget:
    """One slab's fused budget sweep with the halo exchange hidden behind the interior rows.

    The integrals of rows [1+depth, Ny-depth] need no neighbour data, so that sweep runs on the compute stream
    while the ring exchange (pack, NCCL send/recv, unpack) runs on a side stream; the two `depth`-row strips next to
    the slab edges are swept afterwards and the three partial integral vectors are added before the all-reduce."""

    def __init__(self, grid, fields, make_ops, depth=2, rank=None, world=None, group=None):
        from .operations import FusedDiagnostics, Integral
        self.grid, self.group = grid, group
        self.exchanger = SlabHaloExchanger(grid, fields, depth=depth, rank=rank, world=world, group=group)
        Ny = grid.N[1]
        if Ny < 4 * depth:
            raise ValueError("slab too thin to overlap the exchange")
        ops = make_ops()
        win = lambda lo, hi: FusedDiagnostics(*[Integral(o, indices=(None, (lo, hi), None)) for o in ops])
        self.interior = win(1 + depth, Ny - depth)
        self.low, self.high = win(1, depth), win(Ny - depth + 1, Ny)
        self.comm_stream = torch.cuda.Stream(device=grid.device)
        self.done = torch.cuda.Event()
        self.total = torch.zeros(len(ops), dtype=torch.float64, device=grid.device)
        self.n = len(ops)

    def step(self):
        cur = torch.cuda.current_stream(self.grid.device)
        self.comm_stream.wait_stream(cur)
        with torch.cuda.stream(self.comm_stream):
            self.exchanger.exchange()
            self.done.record(self.comm_stream)
        self.interior.plan.execute()                      # overlaps the exchange
        cur.wait_event(self.done)
        self.low.plan.execute()
        self.high.plan.execute()
        torch.add(self.interior.plan.integrals[:self.n], self.low.plan.integrals[:self.n], out=self.total)
        self.total.add_(self.high.plan.integrals[:self.n])
        allreduce_integrals(self.total, self.group)
        return self.total
