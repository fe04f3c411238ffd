"""Fields on the staggered grid: device-resident parent arrays laid out exactly as Oceananigans lays out
a Field's parent (column-major, i fastest, `halo` cells either side), so the same `ocb_field`
descriptor serves the Julia host (CuArray parents) and this Python host (torch CUDA tensors).

A torch tensor of shape (sz, sy, sx), C-contiguous, *is* a column-major (sx, sy, sz) array.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import _lib
from .grids import Center, Face, Bounded, Periodic, RectilinearGrid

_LOC_CODE = {Center: _lib.OCB_CENTER, Face: _lib.OCB_FACE, None: _lib.OCB_NOTHING}


def current_stream_ptr(device):
    if torch.device(device).type != "cuda":
        return None
    return torch.cuda.current_stream(device).cuda_stream


def normalize_location(loc):
    out = []
    for l in loc:
        if l is None:
            out.append(None)
        else:
            s = str(getattr(l, "__name__", l)).lower()
            out.append(Center if s.startswith("c") else Face)
    return tuple(out)


class AbstractField:
    """Anything that can be an operand: has a `grid` (or None) and a `location`."""
    grid = None
    location = (Center, Center, Center)

    def __sub__(self, other):
        return BinaryOperation("-", self, other)

    def __add__(self, other):
        return BinaryOperation("+", self, other)

    def __mul__(self, other):
        return BinaryOperation("*", self, other)

    def __rmul__(self, other):
        return BinaryOperation("*", other, self)


class ZeroField(AbstractField):
    """Oceananigans' ZeroField: reads as 0 everywhere (e.g. `w` of a hydrostatic model, src/Oceanostics.jl:160-162)."""

    def descriptor(self):
        f = _lib.ocb_field()
        f.kind = _lib.OCB_FIELD_ZERO
        return f

    def __repr__(self):
        return "ZeroField"


class ConstantField(AbstractField):
    """A `Number` used where a field is expected (e.g. `U = 0`, test/test_turbulent_kinetic_energy_equation.jl:113)."""

    def __init__(self, value):
        self.value = float(value)

    def descriptor(self):
        f = _lib.ocb_field()
        f.kind = _lib.OCB_FIELD_CONSTANT
        f.value = self.value
        return f

    def __repr__(self):
        return f"ConstantField({self.value})"


def as_operand(x):
    if x is None:
        return ZeroField()
    if isinstance(x, (int, float, np.floating, np.integer)):
        return ConstantField(x) if x != 0 else ZeroField()
    return x


class Field(AbstractField):
    """`Field{LX,LY,LZ}(grid)`; `Field(op)` wraps an operation as a computed field (see operations.py)."""

    def __new__(cls, *args, **kwargs):
        # Field(operation) -> computed field, as in Oceananigans
        if args and not isinstance(args[0], (tuple, list, str)) and hasattr(args[0], "_as_computed_field"):
            return args[0]._as_computed_field(**kwargs)
        return super().__new__(cls)

    def __init__(self, location, grid: RectilinearGrid = None, data: torch.Tensor = None, indices=None):
        if not isinstance(location, (tuple, list, str)):
            return   # Field(op): already built by __new__
        self.grid = grid
        self.location = normalize_location(location)
        self.extent = grid.extent_of(self.location)
        self.halo = tuple(0 if self.location[d] is None else grid.H[d] for d in range(3))
        # `indices`: 1-based inclusive windows per axis (Oceananigans `indices=`); None = whole axis
        self.window = [(1, self.extent[d]) for d in range(3)]
        if indices is not None:
            for d, w in enumerate(indices):
                if w is not None and w != slice(None):
                    lo, hi = (w.start, w.stop) if isinstance(w, slice) else w
                    self.window[d] = (int(lo), int(hi))
        self.windowed = any(self.window[d] != (1, self.extent[d]) for d in range(3))
        stored = tuple(self.window[d][1] - self.window[d][0] + 1 for d in range(3))
        self.stored = stored
        shape = tuple(stored[d] + 2 * self.halo[d] for d in (2, 1, 0))
        if data is None:
            data = torch.zeros(shape, dtype=grid.dtype, device=grid.device)
        if tuple(data.shape) != shape or not data.is_contiguous():
            raise ValueError(f"parent array must be contiguous with shape {shape}; got {tuple(data.shape)}")
        self.data = data
        self.status_time = None
        self.boundary = dict(impenetrable=True, zgrad_bottom=math.nan, zgrad_top=math.nan)

    # ---- views ------------------------------------------------------------------------------------------
    @property
    def interior(self) -> torch.Tensor:
        """View of the stored interior, indexed [k, j, i] (torch order)."""
        hz, hy, hx = self.halo[2], self.halo[1], self.halo[0]
        return self.data[hz:hz + self.stored[2], hy:hy + self.stored[1], hx:hx + self.stored[0]]

    def interior_ijk(self) -> np.ndarray:
        """Host copy of the interior indexed [i, j, k] (Julia order)."""
        return self.interior.detach().cpu().numpy().transpose(2, 1, 0)

    def set(self, value):
        """`set!(field, value)`: a number, an array indexed [i, j, k] of the interior shape, or f(x, y, z)."""
        g = self.grid
        if callable(value):
            idx = [np.arange(self.window[d][0], self.window[d][1] + 1) for d in range(3)]
            x = g.nodes(0, self.location[0] or Center, idx[0]).astype(np.float64).reshape(-1, 1, 1)
            y = g.nodes(1, self.location[1] or Center, idx[1]).astype(np.float64).reshape(1, -1, 1)
            z = g.nodes(2, self.location[2] or Center, idx[2]).astype(np.float64).reshape(1, 1, -1)
            vals = np.vectorize(value, otypes=[np.float64])(*np.broadcast_arrays(x, y, z))
            value = vals
        if isinstance(value, np.ndarray):
            value = torch.from_numpy(np.ascontiguousarray(value.transpose(2, 1, 0))).to(self.data.dtype)
        if isinstance(value, torch.Tensor):
            self.interior.copy_(value.to(self.data.device, self.data.dtype))
        else:
            self.interior.fill_(float(value))
        return self

    # ---- descriptors --------------------------------------------------------------------------------------
    def descriptor(self) -> _lib.ocb_field:
        f = _lib.ocb_field()
        f.data = self.data.data_ptr()
        f.kind = _lib.OCB_FIELD_ARRAY
        strides = (self.data.stride(2), self.data.stride(1), self.data.stride(0))
        for d in range(3):
            f.loc[d] = _LOC_CODE[self.location[d]]
            f.halo[d] = self.halo[d]
            f.size[d] = self.stored[d] + 2 * self.halo[d]
            f.stride[d] = 0 if self.location[d] is None else strides[d]
        return f

    # ---- halos ----------------------------------------------------------------------------------------------
    def fill_halo_regions(self):
        """`fill_halo_regions!` with the field's (default or gradient) boundary conditions -- CUDA only."""
        if self.windowed:
            return self
        if self.data.device.type != "cuda":
            raise _lib.OcbError("fill_halo_regions needs a CUDA field: liboceanostics_b200 has no CPU fallback")
        L = _lib.lib()
        desc, gd = self.descriptor(), self.grid.descriptor()
        import ctypes as C
        _lib.check(L.ocb_fill_halo(C.byref(gd), C.byref(desc), int(self.boundary["impenetrable"]),
                                   float(self.boundary["zgrad_bottom"]), float(self.boundary["zgrad_top"]),
                                   current_stream_ptr(self.data.device)))
        return self

    def fill_halo_regions_xz(self):
        """Halo fill in x and z only: the y halos of a y-slab come from the ring neighbours (distributed.py)."""
        if self.data.device.type != "cuda":
            raise _lib.OcbError("fill_halo_regions needs a CUDA field: liboceanostics_b200 has no CPU fallback")
        import ctypes as C
        desc, gd = self.descriptor(), self.grid.descriptor()
        desc.loc[1] = _lib.OCB_NOTHING          # skip the y pass; x and z passes still run over whole rows
        _lib.check(_lib.lib().ocb_fill_halo(C.byref(gd), C.byref(desc), int(self.boundary["impenetrable"]),
                                            float(self.boundary["zgrad_bottom"]), float(self.boundary["zgrad_top"]),
                                            current_stream_ptr(self.data.device)))
        return self

    def __repr__(self):
        loc = ", ".join({Center: "Center", Face: "Face", None: "Nothing"}[l] for l in self.location)
        return f"{self.extent[0]}×{self.extent[1]}×{self.extent[2]} Field at ({loc}) on {self.grid.device}"


def CenterField(grid, **kw):
    return Field((Center, Center, Center), grid, **kw)


def XFaceField(grid, **kw):
    return Field((Face, Center, Center), grid, **kw)


def YFaceField(grid, **kw):
    return Field((Center, Face, Center), grid, **kw)


def ZFaceField(grid, **kw):
    return Field((Center, Center, Face), grid, **kw)


class BinaryOperation(AbstractField):
    """Lazy `a ∘ b` (Oceananigans BinaryOperation).  The diagnostics planner pattern-matches
    `field - mean` (perturbations, src/TurbulentKineticEnergyEquation.jl:327-330); any other tree is
    materialised with `Field(op)` through the library's pointwise kernel."""

    def __init__(self, op, a, b):
        self.op, self.a, self.b = op, as_operand(a), as_operand(b)
        self.grid = getattr(self.a, "grid", None) or getattr(self.b, "grid", None)
        self.location = getattr(self.a, "location", None) if isinstance(self.a, (Field, BinaryOperation)) else self.b.location

    def _as_computed_field(self, location=None):
        from .operations import materialize_binary
        return materialize_binary(self, location)


def average_xy(field: Field, dims=(1, 2)) -> Field:
    """`Field(Average(field, dims=dims))` of a plain field on regularly spaced reduced axes: a reduced
    Field (e.g. U(z) = ⟨u⟩ₓᵧ).  This is the *caller's* step (Oceananigans' own Average), done with torch."""
    g = field.grid
    loc = tuple(None if (d + 1) in dims else field.location[d] for d in range(3))
    out = Field(loc, g)
    tdims = tuple(2 - (d - 1) for d in dims)      # torch axis of Julia dim d
    out.interior.copy_(field.interior.mean(dim=tdims, keepdim=True))
    out.boundary["impenetrable"] = False
    if out.data.device.type == "cuda":
        out.fill_halo_regions()
    return out
