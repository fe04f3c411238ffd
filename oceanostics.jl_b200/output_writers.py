"""Output side of the path: what an Oceananigans `OutputWriter` does with the diagnostics each
output step -- `compute_at!` every output at the model time, keep only the `indices=` window, optionally average over a time
window, write.  The reference relies on this contract in two places: windowed output fields must compute through the staged
path without writing out of bounds (test/test_spatial_filters.jl:207-222) and a diagnostic recomputed at a new time -- "as an
OutputWriter does each output" -- must see the evolved state all the way through nested Fields
(test/test_subfilter_kinetic_energy_equation.jl:87).

What is new here is the fusion: outputs that are KernelFunctionOperations over compatible operands are grouped into ONE
`FusedDiagnostics` sweep per output step (the sibling registry of SURVEY.md section 8b), windows included, instead of one
launch per output.  Containers: `NetCDFWriter` writes classic NetCDF-3 through scipy (Julia's NCDatasets / Python's netCDF4
are not in the image), `JLD2Writer` a directory-free `.npz` series with the same naming (JLD2 is Julia-only).  Time stepping
itself stays in Oceananigans: `writer(model)` is called by whoever advances `model.clock`.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from .fields import Field
from .operations import (AbstractOperation, ComputedField, FusedDiagnostics, KernelFunctionOperation, ReducedField, Reduction, compute_at)


# ---- schedules (Oceananigans.Utils) ------------------------------------------------------------------------------------
class IterationInterval:
    def __init__(self, interval, offset=0):
        self.interval, self.offset = int(interval), int(offset)

    def __call__(self, clock):
        return (int(clock.iteration) - self.offset) % self.interval == 0


class TimeInterval:
    """Actuates whenever the model time crosses a multiple of `interval` (and at the first call)."""

    def __init__(self, interval):
        self.interval, self.next = float(interval), 0.0

    def __call__(self, clock):
        if clock.time >= self.next - 1e-12 * max(1.0, abs(self.next)):
            self.next = (np.floor(clock.time / self.interval + 1e-9) + 1) * self.interval
            return True
        return False


class AveragedTimeInterval:
    """`AveragedTimeInterval(interval; window)`: every `interval`, write the time average over the preceding `window`
    (Oceananigans' WindowedTimeAverage: a trapezoid-free running sum of field x dt over the steps inside the window)."""

    def __init__(self, interval, window=None):
        self.interval = float(interval)
        self.window = float(window if window is not None else interval)
        if not 0 < self.window <= self.interval:
            raise ValueError("the averaging window must be positive and no longer than the output interval")
        self.next = self.interval

    def collecting(self, time):
        return time > self.next - self.window + 1e-12 * max(1.0, abs(self.next))

    def due(self, time):
        return time >= self.next - 1e-12 * max(1.0, abs(self.next))

    def advance(self):
        self.next += self.interval


# ---- outputs -------------------------------------------------------------------------------------------------------------
def _normalize_indices(indices):
    """Oceananigans `indices=(:, :, Nz)` style -> per-axis None or 1-based inclusive (lo, hi)."""
    if indices is None:
        return None
    out = []
    for w in indices:
        if w is None or w == slice(None) or w == ":":
            out.append(None)
        elif isinstance(w, int):
            out.append((w, w))
        elif isinstance(w, slice):
            out.append((w.start, w.stop))
        else:
            out.append((int(w[0]), int(w[1])))
    return tuple(out)


class _Output:
    def __init__(self, name, obj, indices):
        self.name = name
        if isinstance(obj, Reduction):
            self.field, self.kind = obj._as_computed_field(), "scalar"
        elif isinstance(obj, ReducedField):
            self.field, self.kind = obj, "scalar"
        elif isinstance(obj, KernelFunctionOperation):
            self.field, self.kind = ComputedField(obj, indices=indices), "field"       # the window is computed, nothing else
        elif isinstance(obj, AbstractOperation) and not isinstance(obj, Field):
            self.field, self.kind = Field(obj), "field"
        elif isinstance(obj, Field):
            self.field, self.kind = obj, "field"
        else:
            raise TypeError(f"output `{name}` is neither a Field nor an operation")
        self.indices = indices
        self.acc = None            # running sum of value x dt on the device (time-averaged outputs)

    def values(self):
        """The output's current values on the device: the stored interior of a (windowed) field, cut to `indices` when the
        field itself is a full one; the scalar of a reduction."""
        f = self.field
        if self.kind == "scalar":
            v = f.result.clone()
            if f.reduction.kind == "average":
                v /= f.grid.total_volume(f.operand.location)
            return v
        it = f.interior
        if self.indices is not None and not getattr(f, "windowed", False):
            sl = []
            for d in (2, 1, 0):                       # torch order (k, j, i)
                w = self.indices[d]
                sl.append(slice(None) if w is None else slice(w[0] - 1, w[1]))
            it = it[tuple(sl)]
        return it


class _Writer:
    """Common machinery of the two writers."""

    def __init__(self, model, outputs, filename, schedule, indices=None, overwrite_existing=True, dir="."):
        self.model, self.schedule = model, schedule
        self.indices = _normalize_indices(indices)
        self.outputs = [_Output(name, obj, self.indices) for name, obj in dict(outputs).items()]
        self.path = os.path.join(dir, filename)
        if os.path.exists(self.path) and not overwrite_existing:
            raise FileExistsError(self.path)
        self.times, self.records = [], {o.name: [] for o in self.outputs}
        # siblings: every computed KFO output / reduction whose operands agree shares one sweep per output step
        self.groups, loose = [], []
        pending = [o for o in self.outputs if isinstance(o.field, (ComputedField, ReducedField)) and getattr(o.field, "group", None) is None
                   and hasattr(getattr(o.field, "operand", None), "operands")]
        while pending:
            head, rest, members = pending[0], [], [pending[0]]
            merged = head.field.operand.operands
            for o in pending[1:]:
                m2 = merged.merged_with(o.field.operand.operands)
                if m2 is not None and len(members) < 16 and o.field.grid is head.field.grid:
                    members.append(o); merged = m2
                else:
                    rest.append(o)
            if len(members) > 1:
                self.groups.append(FusedDiagnostics(*[m.field for m in members]))
            else:
                loose.append(head)
            pending = rest
        self._last_time = None
        self._window_elapsed = 0.0

    # -- one output step ---------------------------------------------------------------------------------------------
    def _compute_all(self, time):
        for g in self.groups:
            g.compute(time)
        for o in self.outputs:
            if getattr(o.field, "group", None) is not None:
                continue                                          # computed with its siblings above
            if getattr(o.field, "is_computed", False) or isinstance(o.field, ReducedField):
                compute_at(o.field, time)

    def __call__(self, model=None):
        """Call after every model step (Oceananigans calls its writers from `run!`): collects / writes when the schedule says so."""
        clock = (model or self.model).clock
        time = float(clock.time)
        s = self.schedule
        if isinstance(s, AveragedTimeInterval):
            if self._last_time is None:            # first call: the next output is the first multiple of the interval after now
                s.next = (np.floor(time / s.interval + 1e-9) + 1) * s.interval
            dt = 0.0 if self._last_time is None else time - self._last_time
            self._last_time = time
            if s.collecting(time) and dt > 0:
                self._compute_all(time)
                for o in self.outputs:
                    v = o.values()
                    if o.acc is None:
                        o.acc = torch.zeros_like(v, dtype=torch.float64)
                    o.acc.add_(v, alpha=dt)                       # running integral of the output over the window
                self._window_elapsed += dt
            if s.due(time) and self._window_elapsed > 0:
                self._append(time, [(o.acc / self._window_elapsed) for o in self.outputs])
                for o in self.outputs:
                    o.acc.zero_()
                self._window_elapsed = 0.0
                s.advance()
                return True
            return False
        if not s(clock):
            return False
        self._compute_all(time)
        self._append(time, [o.values() for o in self.outputs])
        return True

    def _append(self, time, values):
        self.times.append(time)
        for o, v in zip(self.outputs, values):
            a = v.detach().to("cpu").numpy()
            self.records[o.name].append(np.array(a, copy=True).reshape(()) if o.kind == "scalar" else np.array(a, copy=True))
        self.flush()

    def flush(self):
        raise NotImplementedError


class NetCDFWriter(_Writer):
    """`NetCDFWriter(model, outputs; filename, schedule, indices)`: classic NetCDF-3 (scipy.io.netcdf_file); variables are
    (time, z, y, x) for fields -- the window's extents -- and (time,) for reductions."""

    def flush(self):
        from scipy.io import netcdf_file
        with netcdf_file(self.path, "w", version=2) as nc:
            nc.createDimension("time", len(self.times))
            t = nc.createVariable("time", "f8", ("time",))
            t[:] = np.asarray(self.times)
            for o in self.outputs:
                recs = self.records[o.name]
                if o.kind == "scalar":
                    var = nc.createVariable(o.name, "f8", ("time",))
                    var[:] = np.asarray(recs, dtype=np.float64)
                    continue
                shp = recs[0].shape
                dims = []
                for ax, n in zip("zyx", shp):
                    dname = f"{ax}_{o.name}" if nc.dimensions.get(ax, n) != n else ax
                    if dname not in nc.dimensions:
                        nc.createDimension(dname, n)
                    dims.append(dname)
                var = nc.createVariable(o.name, "f8" if recs[0].dtype == np.float64 else "f4", ("time",) + tuple(dims))
                var[:] = np.stack(recs)
                var.location = "".join(o.field.location[d] or "n" for d in range(3)).encode()


class JLD2Writer(_Writer):
    """`JLD2Writer(model, outputs; filename, schedule, indices)` with an `.npz` container: `timeseries/<name>/<n>` arrays
    and `timeseries/t/<n>` times, the JLD2 layout Oceananigans writes, flattened into npz keys."""

    def flush(self):
        out = {}
        for n, t in enumerate(self.times):
            out[f"timeseries/t/{n}"] = np.asarray(t)
            for o in self.outputs:
                out[f"timeseries/{o.name}/{n}"] = self.records[o.name][n]
        path = self.path if self.path.endswith(".npz") else self.path + ".npz"
        np.savez(path, **out)
