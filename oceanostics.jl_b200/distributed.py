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


_IPC_MAPPED = {}     # (device index, 64-byte handle) -> mapped base address: a handle is opened once per process


def _ipc_export(tensor):
    handle, off = (C.c_ubyte * 64)(), C.c_int64()
    with torch.cuda.device(tensor.device):
        _lib.check(_lib.lib().ocb_ipc_export(tensor.data_ptr(), handle, C.byref(off)))
    return bytes(handle), off.value


def _ipc_map(device, handle, offset):
    key = (torch.device(device).index, handle)
    if key not in _IPC_MAPPED:
        base = C.c_void_p()
        with torch.cuda.device(device):
            _lib.check(_lib.lib().ocb_ipc_import((C.c_ubyte * 64)(*handle), C.byref(base)))
        _IPC_MAPPED[key] = base.value
    return _IPC_MAPPED[key] + offset


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


def _pack_many_cuda(fields, h, buffers, unpack):
    """Both edges of every field in one launch (ocb_halo_pack_y_many): buffers[2*f + side]."""
    L = _lib.lib()
    descs = (_lib.ocb_field * len(fields))(*[f.descriptor() for f in fields])
    ptrs = (C.c_void_p * len(buffers))(*[b.data_ptr() for b in buffers])
    dt = _lib.OCB_F64 if fields[0].data.dtype == torch.float64 else _lib.OCB_F32
    _lib.check(L.ocb_halo_pack_y_many(dt, descs, len(fields), h, ptrs, 1 if unpack else 0, current_stream_ptr(fields[0].data.device)))


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
        batched = self.pack is _pack_cuda and self.unpack is _unpack_cuda and len(self.fields) <= 8 and \
            len({f.data.dtype for f in self.fields}) == 1
        if batched:      # one launch for all fields and both edges
            _pack_many_cuda(self.fields, h, [b for pair in zip(self.send_lo, self.send_hi) for b in pair], unpack=False)
        else:
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
        if batched:
            _pack_many_cuda(self.fields, h, [b for pair in zip(self.recv_lo, self.recv_hi) for b in pair], unpack=True)
        else:
            for f, rlo, rhi in zip(self.fields, self.recv_lo, self.recv_hi):
                self.unpack(f, 0, h, rlo)
                self.unpack(f, 1, h, rhi)


class PeerSlabExchanger:
    """The same ring exchange by direct stores into the neighbours' halo rows over NVLink (CUDA IPC, one process per GPU of
    one box): ONE kernel per step instead of pack + NCCL send/recv + unpack, then a flag word per neighbour.

    `push()` stores this rank's edge rows into both neighbours and publishes the step number; `wait()` is a stream-ordered
    wait for both neighbours' pushes of the current step.  Ordering contract (include/oceanostics_b200.h): the next push must
    follow a collective that every rank enters after it has read its halos (here: the all-reduce of the integrals)."""

    def __init__(self, grid, fields, depth=2, rank=None, world=None, group=None, timeout=10.0, ring=None):
        self.grid, self.fields, self.h, self.group, self.timeout = grid, list(fields), int(depth), group, float(timeout)
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world = dist.get_world_size(group) if world is None else world
        if self.h < 1 or self.h > grid.H[1]:
            raise ValueError(f"exchange depth {depth} must be within 1..halo ({grid.H[1]})")
        if len(self.fields) > 8 or len({f.data.dtype for f in self.fields}) != 1:
            raise ValueError("the peer exchange takes up to 8 fields of one dtype")
        dev = grid.device
        if dev.type != "cuda":
            raise _lib.OcbError("the peer-memory exchange needs CUDA devices (no CPU fallback)")
        self.lo_rank, self.hi_rank = (self.rank - 1) % self.world, (self.rank + 1) % self.world
        # [from low, from high, timed out, unused]: the first words of the ring's sync block when the exchange is part of one
        self.ring = ring
        self.flags = ring.flags if ring is not None else torch.zeros(4, dtype=torch.int64, device=dev)
        self.step_no = 0
        ny = int(grid.N[1])
        if self.world == 1:
            peers = {self.rank: {"fields": [(f.data.data_ptr(), ny) for f in self.fields], "flags": self.flags.data_ptr()}}
        else:
            torch.cuda.synchronize(dev)
            mine = {"fields": [_ipc_export(f.data) + (ny,) for f in self.fields], "flags": _ipc_export(self.flags)}
            everyone = [None] * self.world
            dist.all_gather_object(everyone, mine, group=group)
            peers = {r: {"fields": [(_ipc_map(dev, hd, off), n) for hd, off, n in everyone[r]["fields"]],
                         "flags": _ipc_map(dev, *everyone[r]["flags"])} for r in {self.lo_rank, self.hi_rank}}
            dist.barrier(group=group)
        lo, hi = peers[self.lo_rank], peers[self.hi_rank]
        nf = len(self.fields)
        self._ptrs = (C.c_void_p * (2 * nf))(*[p for f in range(nf) for p in (lo["fields"][f][0], hi["fields"][f][0])])
        self._nys = (C.c_int32 * (2 * nf))(*[n for f in range(nf) for n in (lo["fields"][f][1], hi["fields"][f][1])])
        self._lo_flag, self._hi_flag = lo["flags"] + 8, hi["flags"]      # I am my low neighbour's high neighbour
        self._descs = (_lib.ocb_field * nf)(*[f.descriptor() for f in self.fields])
        self._dt = _lib.OCB_F64 if self.fields[0].data.dtype == torch.float64 else _lib.OCB_F32
        self.bytes_per_exchange = sum(2 * f.data.shape[0] * self.h * f.data.shape[2] * f.data.element_size() for f in self.fields)

    def push(self):
        self.step_no += 1
        _lib.check(_lib.lib().ocb_halo_push_y_many(self._dt, self._descs, len(self.fields), self.h, self._ptrs, self._nys, self._lo_flag,
                                                   self._hi_flag, self.step_no, current_stream_ptr(self.grid.device)))

    def wait(self):
        _lib.check(_lib.lib().ocb_halo_wait(self.flags.data_ptr(), self.step_no, self.timeout, current_stream_ptr(self.grid.device)))

    def exchange(self):
        self.push()
        self.wait()
        self._unchecked = getattr(self, "_unchecked", 0) + 1
        if self._unchecked >= 64:          # a wait that gave up must not go unnoticed for long (one sync per 64 exchanges)
            self.check()

    def check(self):
        """Raise if a wait gave up (a neighbour never pushed).  Synchronises the device."""
        self._unchecked = 0
        if int(self.flags[2].item()) != 0:
            raise _lib.OcbError("halo exchange: a neighbour's rows did not arrive within the timeout")

    def close(self):
        pass      # mappings are process-wide (_IPC_MAPPED) and live until the process exits


class PeerRing:
    """The sync blocks of a peer-memory ring (`ocb_ring`, include/oceanostics_b200.h): every rank owns one zeroed device
    block of OCB_RING_SYNC_WORDS 64-bit words -- halo flags, timeout word, and the mailboxes of the integral all-reduce -- and
    maps every other rank's over CUDA IPC.  torch.distributed only carries the 64-byte handles at set-up."""

    def __init__(self, device, rank=None, world=None, group=None, timeout=10.0):
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world = dist.get_world_size(group) if world is None else world
        if self.world > _lib.OCB_MAX_RANKS:
            raise ValueError(f"a peer ring holds at most {_lib.OCB_MAX_RANKS} ranks")
        if torch.device(device).type != "cuda":
            raise _lib.OcbError("the peer-memory ring needs CUDA devices (no CPU fallback)")
        self.device, self.group, self.timeout = torch.device(device), group, float(timeout)
        self.block = torch.zeros(_lib.OCB_RING_SYNC_WORDS, dtype=torch.int64, device=self.device)
        ptrs = [None] * self.world
        ptrs[self.rank] = self.block.data_ptr()
        if self.world > 1:
            torch.cuda.synchronize(self.device)
            everyone = [None] * self.world
            dist.all_gather_object(everyone, _ipc_export(self.block), group=group)
            for r, (hd, o) in enumerate(everyone):
                if r != self.rank:
                    ptrs[r] = _ipc_map(self.device, hd, o)
            dist.barrier(group=group)
        self.ptrs = ptrs
        self.desc = _lib.ocb_ring()
        self.desc.rank, self.desc.nranks, self.desc.timeout_seconds = self.rank, self.world, self.timeout
        for r in range(self.world):
            self.desc.sync[r] = ptrs[r]

    @property
    def flags(self):
        """words 0..3 of the own block: halo flag from the low / high neighbour, timed-out word, spare"""
        return self.block[:4]

    def timed_out(self):
        return int(self.block[2].item()) != 0


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

    def __init__(self, grid, fields, make_ops, depth=2, rank=None, world=None, group=None, transport="nccl"):
        from .operations import FusedDiagnostics, Integral
        self.grid, self.group, self.transport = grid, group, transport
        if transport == "peer":        # direct stores into the neighbours' halos; the all-reduce below orders their re-use
            self.exchanger = PeerSlabExchanger(grid, fields, depth=depth, rank=rank, world=world, group=group)
        else:
            self.exchanger = SlabHaloExchanger(grid, fields, depth=depth, rank=rank, world=world, group=group)
        Ny = grid.N[1]
        if Ny < 4 * depth:
            raise ValueError("slab too thin to overlap the exchange")
        ops = make_ops()
        win = lambda lo, hi: FusedDiagnostics(*[Integral(o, indices=(None, (lo, hi), None)) for o in ops])
        self.interior = win(1 + depth, Ny - depth)
        self.low, self.high = win(1, depth), win(Ny - depth + 1, Ny)
        self.comm_stream = torch.cuda.Stream(device=grid.device, priority=-1)   # pack / unpack ahead of the sweep's blocks
        self.done = torch.cuda.Event()
        self.total = torch.zeros(len(ops), dtype=torch.float64, device=grid.device)
        self.n = len(ops)

    def step(self, timeline=None):
        """One overlapped step.  `timeline` (a dict) receives CUDA events marking the start, the end of the exchange (side
        stream), the end of the interior sweep, of the edge strips and of the all-reduce -- see `timeline_ms`."""
        cur = torch.cuda.current_stream(self.grid.device)
        mark = (lambda name, stream: None) if timeline is None else (
            lambda name, stream: timeline.setdefault(name, torch.cuda.Event(enable_timing=True)).record(stream))
        mark("start", cur)
        self.comm_stream.wait_stream(cur)
        with torch.cuda.stream(self.comm_stream):
            if self.transport == "peer":
                self.exchanger.push()
            else:
                self.exchanger.exchange()
            self.done.record(self.comm_stream)
            mark("exchange_end", self.comm_stream)
        self.interior.plan.execute()                      # overlaps the exchange
        mark("interior_end", cur)
        cur.wait_event(self.done)
        if self.transport == "peer":
            self.exchanger.wait()                         # the neighbours' rows of this step have landed
        self.low.plan.execute()
        self.high.plan.execute()
        mark("strips_end", cur)
        torch.add(self.interior.plan.integrals[:self.n], self.low.plan.integrals[:self.n], out=self.total)
        self.total.add_(self.high.plan.integrals[:self.n])
        allreduce_integrals(self.total, self.group)
        mark("end", cur)
        if self.transport == "peer":
            self._unchecked = getattr(self, "_unchecked", 0) + 1
            if self._unchecked >= 64:      # a halo wait that gave up must not go unnoticed (one sync per 64 steps)
                self._unchecked = 0
                self.exchanger.check()
        return self.total

    @staticmethod
    def timeline_ms(timeline):
        """Milliseconds from the start of a step recorded with `step(timeline=...)` (after a device synchronize)."""
        return {k: round(timeline["start"].elapsed_time(v), 4) for k, v in timeline.items() if k != "start"}


class RingSlabBudget:
    """One slab's fused budget step over a peer-memory ring: THREE launches and no NCCL on the data path.

      1. `ocb_halo_push_y_many` (side stream, high priority): ONE row of every input field stored straight into each ring
         neighbour's halo row over NVLink, the step number published behind it.  One row is all the class-sum kernel reads
         beyond its slab (its advection term is summed by parts), half of what the cell-exact kernels need.
      2. the class-sum sweep over the whole slab: the tile rows that read a halo row are last in launch order and wait in the
         kernel for the neighbour's flag, so every other row overlaps the exchange.
      3. one block sums the per-block partials, stores them into every rank's sync block and adds all ranks' in rank order.

    `step()` returns the device vector of GLOBAL integrals (bit-identical on every rank).  A wait that gives up (a neighbour
    that never pushed) turns the integrals into NaN on every rank and sets the ring's timed-out word: never a hang, never a
    silently stale halo."""

    def __init__(self, grid, fields, make_ops, rank=None, world=None, group=None, timeout=10.0):
        from .operations import FusedDiagnostics, Integral
        self.grid, self.group = grid, group
        self.ring = PeerRing(grid.device, rank=rank, world=world, group=group, timeout=timeout)
        self.exchanger = PeerSlabExchanger(grid, fields, depth=1, rank=self.ring.rank, world=self.ring.world, group=group,
                                           timeout=timeout, ring=self.ring)
        self.diag = FusedDiagnostics(*[Integral(o) for o in make_ops()])
        self.plan = self.diag.plan
        if not self.plan.ring_capable:
            raise _lib.OcbError("these diagnostics do not run on the class-sum kernel over a whole y-periodic slab: use "
                                "OverlappedSlabBudget (exchange + sweep + all-reduce) instead")
        self.n = self.plan.n_slots
        self.total = self.plan.integrals[:self.n]
        # slabs of one or two tile rows have no row to hide the exchange behind: push on the compute stream, then sweep
        self.serial = grid.N[1] < 48
        self.comm_stream = torch.cuda.Stream(device=grid.device, priority=-1)

    def step(self, timeline=None):
        cur = torch.cuda.current_stream(self.grid.device)
        mark = (lambda name, stream: None) if timeline is None else (
            lambda name, stream: timeline.setdefault(name, torch.cuda.Event(enable_timing=True)).record(stream))
        mark("start", cur)
        if self.serial:
            self.exchanger.push()
            mark("exchange_end", cur)
        else:
            # the push of this step follows the previous step's reduction (every rank has read its halo rows by then)
            self.comm_stream.wait_stream(cur)
            with torch.cuda.stream(self.comm_stream):
                self.exchanger.push()
                mark("exchange_end", self.comm_stream)
        self.plan.execute_ring(self.ring, self.exchanger.step_no)
        mark("end", cur)
        return self.total

    def check(self):
        if self.ring.timed_out():
            raise _lib.OcbError("ring step: a neighbour's rows or partial integrals did not arrive within the timeout")

    timeline_ms = staticmethod(lambda timeline: {k: round(timeline["start"].elapsed_time(v), 4) for k, v in timeline.items() if k != "start"})


def _axis_spec(grid, d):
    f = grid.nodes(d, "f", list(range(1, grid.N[d] + 2)))
    return (float(f[0]), float(f[-1])) if grid.uniform[d] else f


class SlabFilteredField:
    """`Field(filter(ψ))` for a field that lives on a y slab.

    The x and z passes of the staged filter never leave the slab; the y pass reaches `w` rows into each ring
    neighbour (w = the filter's half width along y, 8 rows in BASELINE config 5).  Those rows of ψ are exchanged once
    (one grouped NCCL send/recv), the slab is filtered on a grid extended by w rows on each cut side with the ordinary
    staged kernels, and the slab's own rows are kept: every retained output sees exactly the taps, weights and
    summation order of the single-domain filter, so the result is bit-identical to it.  With a Bounded y axis the two
    end ranks keep their true boundary, where the filter's boundary policy applies as usual."""

    def __init__(self, psi, make_filter, rank=None, world=None, group=None, periodic_y=True):
        from .fields import Field
        from .grids import RectilinearGrid, Bounded
        from .SpatialFilters import FilterOperation
        self.psi, self.group = psi, group
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world = dist.get_world_size(group) if world is None else world
        g = psi.grid
        probe = make_filter(psi)
        if not isinstance(probe, FilterOperation):
            raise TypeError("make_filter(ψ) must return BoxFilter(ψ; ...) or GaussianFilter(ψ; ...)")
        self.w = probe.widths[probe.sorted_dims.index(2)] if 2 in probe.sorted_dims else 0
        if psi.location[1] != "c":
            raise ValueError("slab filters take fields located at cell centres in y")
        if not g.uniform[1]:
            raise ValueError("slab filters need a regularly spaced y axis")
        ny = g.N[1]
        if self.w > ny:
            raise ValueError(f"filter half width {self.w} along y exceeds the slab's {ny} rows")
        first, last = self.rank == 0, self.rank == self.world - 1
        self.ext_lo = self.w if (periodic_y or not first) else 0          # rows borrowed from the low / high neighbour
        self.ext_hi = self.w if (periodic_y or not last) else 0
        yf = g.nodes(1, "f", [1, ny + 1])
        dy = float(yf[1] - yf[0]) / ny
        topo = (g.user_topology[0], Bounded, g.user_topology[2])
        self.ext_grid = RectilinearGrid((g.N[0], ny + self.ext_lo + self.ext_hi, g.N[2]), x=_axis_spec(g, 0),
                                        y=(float(yf[0]) - self.ext_lo * dy, float(yf[1]) + self.ext_hi * dy), z=_axis_spec(g, 2),
                                        halo=g.H, topology=topo, dtype=g.dtype, device=g.device)
        self.ext_psi = Field(psi.location, self.ext_grid)
        self.ext_filtered = Field(make_filter(self.ext_psi))
        self.result = Field(psi.location, g)
        self.lo_rank, self.hi_rank = (self.rank - 1) % self.world, (self.rank + 1) % self.world
        shp = lambda n: (psi.interior.shape[0], n, psi.interior.shape[2])
        mk = lambda n: torch.empty(shp(n), dtype=psi.data.dtype, device=psi.data.device)
        self.send_lo, self.send_hi = mk(self.w), mk(self.w)               # my first / last w rows
        self.recv_lo, self.recv_hi = mk(self.ext_lo), mk(self.ext_hi)
        self.bytes_per_exchange = (self.send_lo.numel() + self.send_hi.numel()) * self.send_lo.element_size()

    def compute(self, time=None):
        from .operations import compute_at
        if getattr(self.psi, "is_computed", False):
            compute_at(self.psi, time)
        self.exchange()
        ny = self.psi.grid.N[1]
        self.ext_filtered.compute(time)
        self.result.interior.copy_(self.ext_filtered.interior[:, self.ext_lo:self.ext_lo + ny, :])
        return self.result

    def exchange(self):
        """Fill the extended copy of ψ: the slab's own rows plus w rows from each ring neighbour."""
        psi, w, ny = self.psi, self.w, self.psi.grid.N[1]
        src, ext = psi.interior, self.ext_psi.interior
        ext[:, self.ext_lo:self.ext_lo + ny, :].copy_(src)
        if w > 0:
            self.send_lo.copy_(src[:, :w, :])
            self.send_hi.copy_(src[:, ny - w:, :])
            if self.world == 1:
                if self.ext_hi: self.recv_hi.copy_(self.send_lo)
                if self.ext_lo: self.recv_lo.copy_(self.send_hi)
            else:
                ops = []
                # my low rows extend the low neighbour's slab upwards, my high rows the high neighbour's downwards
                # (in a ring of two both neighbours are the same peer and messages pair up in posting order: sends low
                # then high, receives high then low, as in SlabHaloExchanger)
                if self.ext_lo: ops.append(dist.P2POp(dist.isend, self.send_lo, self.lo_rank, self.group))
                if self.ext_hi: ops.append(dist.P2POp(dist.isend, self.send_hi, self.hi_rank, self.group))
                if self.ext_hi: ops.append(dist.P2POp(dist.irecv, self.recv_hi, self.hi_rank, self.group))
                if self.ext_lo: ops.append(dist.P2POp(dist.irecv, self.recv_lo, self.lo_rank, self.group))
                for req in dist.batch_isend_irecv(ops):
                    req.wait()
            if self.ext_lo: ext[:, :self.ext_lo, :].copy_(self.recv_lo)
            if self.ext_hi: ext[:, self.ext_lo + ny:, :].copy_(self.recv_hi)
