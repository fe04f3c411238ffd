"""SpatialFilters: separable box / Gaussian low-pass filters (reference: src/SpatialFilters/SpatialFilters.jl,
box_filter.jl, gaussian_filter.jl).  Same constructors, keyword arguments, validation rules and error messages;
`Field(filter(ψ))` takes the staged path (d sequential 1-D passes, _compute_staged_filter!, SpatialFilters.jl:347-378)
on the GPU through `ocb_filter_execute`.

    BoxFilter(ψ; dims, N, boundary="shrink")          GaussianFilter(ψ; dims, σ, N=None, boundary="shrink")
    BoxFilter(; dims, N, boundary)  -> operator       GaussianFilter(; dims, σ, N, boundary) -> operator

`boundary` is "shrink", "edge", a dict(left=a, right=b) (constant padding), or a tuple with one entry per dim in
`dims`; periodic directions always wrap (SpatialFilters.jl:193-197).
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch

from . import _lib
from .fields import AbstractField, Field, current_stream_ptr
from .grids import Periodic
from .operations import AbstractOperation, KernelFunctionOperation, compute_at


# ---- validation (SpatialFilters.jl:234-271) ------------------------------------------------------------------
def tuplefy_dims(dims):
    return (int(dims),) if isinstance(dims, (int, np.integer)) and not isinstance(dims, bool) else dims


def validate_dims(dims):
    if not isinstance(dims, tuple):
        raise ValueError(f"`dims` must be a tuple of integers; got {type(dims).__name__}")
    ok = (len(dims) > 0 and all(isinstance(d, (int, np.integer)) and d in (1, 2, 3) for d in dims) and len(set(dims)) == len(dims))
    if not ok:
        raise ValueError(f"`dims` must be a non-empty tuple of distinct integers drawn from (1, 2, 3); got {dims}")


def validate_N(N):
    if not isinstance(N, (int, np.integer)) or isinstance(N, bool):
        raise ValueError(f"`N` must be an odd integer ≥ 3; got {type(N).__name__}")
    if not (N >= 3 and N % 2 == 1):
        raise ValueError(f"`N` must be an odd integer ≥ 3; got {N}")


def parse_boundary_spec(s):
    if isinstance(s, str):
        if s in ("shrink", "edge"):
            return (s,)
        raise ValueError(f"`boundary` symbol must be :shrink or :edge; got :{s}")
    if isinstance(s, dict):
        if set(s.keys()) != {"left", "right"}:
            raise ValueError(f"`boundary` NamedTuple must have exactly keys `:left` and `:right`; got keys {tuple(s.keys())}")
        return ("constant", float(s["left"]), float(s["right"]))
    raise ValueError(f"`boundary` must be :shrink, :edge, or (left=a, right=b); got {s!r}")


def resolve_filter_policies(grid, extent, dims, boundary):
    """-> (sorted dims, policies); policy = (kind, [left, right], extent along the dim) (SpatialFilters.jl:166-202)."""
    validate_dims(dims)
    if isinstance(boundary, tuple):
        if len(boundary) != len(dims):
            raise ValueError("`boundary` must be a single spec or a tuple with one entry per dim in `dims`; "
                             f"got length {len(boundary)} for dims={dims}")
        specs = boundary
    else:
        specs = tuple(boundary for _ in dims)
    for s in specs:
        parse_boundary_spec(s)
    sorted_dims = tuple(d for d in (1, 2, 3) if d in dims)
    policies = []
    for d in sorted_dims:
        base = ("periodic",) if grid.topology[d - 1] == Periodic else parse_boundary_spec(specs[dims.index(d)])
        policies.append(base + (extent[d - 1],))
    return sorted_dims, tuple(policies)


def validate_periodic_widths(grid, sorted_dims, policies, widths):
    for d, pol, w in zip(sorted_dims, policies, widths):
        if pol[0] == "periodic" and 2 * w + 1 > 2 * grid.N[d - 1] + 1:
            raise ValueError(f"for the periodic direction d={d}, `N` ({2 * w + 1}) exceeds the maximum allowed "
                             f"2*Nd_grid+1 = {2 * grid.N[d - 1] + 1}")


def gaussian_weights(width, sigma_cells, np_dtype):
    """gaussian_filter.jl:312, evaluated in the grid's float type."""
    ft = np.dtype(np_dtype).type
    sig = ft(sigma_cells)
    return [ft(np.exp(-ft((idx - width - 1) ** 2) / (ft(2) * sig * sig))) for idx in range(1, 2 * width + 2)]


def infer_width(sigma, grid, d):
    """gaussian_filter.jl:488: ceil(2σ / Δ_min)."""
    sp = grid.spacings(d - 1, "c", np.arange(1, grid.N[d - 1] + 1)).astype(np.float64)
    return int(math.ceil(2 * sigma / float(sp.min())))


def resolve_gaussian_widths(N, sigma, grid, dims, sorted_dims):
    if N is None:
        return tuple(infer_width(sigma, grid, d) for d in sorted_dims)
    if isinstance(N, tuple):
        if len(N) != len(dims):
            raise ValueError(f"`N` tuple must have one entry per dim in `dims`; got length {len(N)} for dims={dims}")
        for n in N:
            validate_N(n)
        return tuple((N[dims.index(d)] - 1) // 2 for d in sorted_dims)
    validate_N(N)
    return tuple((N - 1) // 2 for _ in sorted_dims)


_POLICY_CODE = {"periodic": _lib.OCB_POLICY_PERIODIC, "shrink": _lib.OCB_POLICY_SHRINK, "edge": _lib.OCB_POLICY_EDGE,
                "constant": _lib.OCB_POLICY_CONSTANT}


# ---- the filter operation and its computed field ------------------------------------------------------------------
class FilterOperation(AbstractOperation):
    """`BoxFilter(ψ; ...)` / `GaussianFilter(ψ; ...)`: an operation located where ψ is, with ψ's own extent."""

    def __init__(self, kind, operand, dims, N=None, sigma=None, boundary="shrink"):
        if isinstance(operand, KernelFunctionOperation) or (isinstance(operand, AbstractOperation) and not isinstance(operand, Field)):
            operand = Field(operand)            # the staged path evaluates the operand once into a Field
        if not isinstance(operand, Field):
            raise TypeError("a spatial filter takes a Field or an operation of this library")
        self.kind, self.operand, self.grid = kind, operand, operand.grid
        self.location = operand.location
        grid = self.grid
        self.dims = tuplefy_dims(dims)
        extent = operand.extent
        if kind == "box":
            validate_N(N)
            self.sorted_dims, self.policies = resolve_filter_policies(grid, extent, self.dims, boundary)
            self.widths = tuple((N - 1) // 2 for _ in self.sorted_dims)
        else:
            if not isinstance(sigma, (int, float, np.floating)) or isinstance(sigma, bool) or not sigma > 0:
                raise ValueError(f"`σ` must be a positive number; got {sigma!r}")
            self.sorted_dims, self.policies = resolve_filter_policies(grid, extent, self.dims, boundary)
            self.widths = resolve_gaussian_widths(N, sigma, grid, self.dims, self.sorted_dims)
        validate_periodic_widths(grid, self.sorted_dims, self.policies, self.widths)
        for w in self.widths:
            if w > _lib.OCB_MAX_FILTER_WIDTH:
                raise _lib.OcbError(f"filter half-width {w} exceeds the library maximum {_lib.OCB_MAX_FILTER_WIDTH}")
        self.sigma = None if sigma is None else float(sigma)
        self.name = "BoxFilter" if kind == "box" else "GaussianFilter"

    def _as_computed_field(self, indices=None):
        return FilteredField(self, indices=indices)

    def build_passes(self):
        """The `ocb_filter_pass` array of the staged evaluation (sorted dims: x, then y, then z)."""
        grid, loc = self.grid, self.location
        passes = (_lib.ocb_filter_pass * len(self.sorted_dims))()
        keep = []
        for n, (d, pol, w) in enumerate(zip(self.sorted_dims, self.policies, self.widths)):
            p = passes[n]
            p.dim, p.width, p.policy, p.extent = d, w, _POLICY_CODE[pol[0]], pol[-1]
            if pol[0] == "constant":
                p.left, p.right = pol[1], pol[2]
            if self.kind == "box":
                p.kind = _lib.OCB_FILTER_BOX
                continue
            sp = grid.spacings(d - 1, "c", np.arange(1, grid.N[d - 1] + 1))
            if sp.min() == sp.max():                      # direction_is_uniform (gaussian_filter.jl:337-340)
                p.kind = _lib.OCB_FILTER_GAUSSIAN
                ft = grid.np_dtype
                sigma_cells = ft(self.sigma) / ft(sp.min())
                for t, wt in enumerate(gaussian_weights(w, sigma_cells, ft)):
                    p.weights[t] = float(wt)
            else:
                p.kind = _lib.OCB_FILTER_GAUSSIAN_STRETCHED
                idx = np.arange(1, pol[-1] + 1)
                nodes = torch.from_numpy(np.ascontiguousarray(grid.nodes(d - 1, loc[d - 1], idx))).to(grid.device)
                spac = torch.from_numpy(np.ascontiguousarray(grid.spacings(d - 1, loc[d - 1], idx))).to(grid.device)
                keep += [nodes, spac]
                p.sigma, p.period = self.sigma, grid.L[d - 1]
                p.nodes, p.spacings = nodes.data_ptr(), spac.data_ptr()
        return passes, keep


class FilteredField(Field):
    """`Field(filter(ψ))`: computed by the staged separable path."""
    is_computed = True

    def __new__(cls, *a, **k):
        return object.__new__(cls)

    def __init__(self, operand: FilterOperation, indices=None):
        if getattr(self, "_initialized", False):
            return
        self._initialized = True
        Field.__init__(self, operand.location, operand.grid, indices=indices)
        self.operand = operand
        self.group = None
        self.boundary["impenetrable"] = False
        self._passes, self._keep = operand.build_passes()

    def compute(self, time=None):
        op = self.operand
        grid = self.grid
        if grid.device.type != "cuda":
            raise _lib.OcbError("filters are evaluated by the CUDA library only (no CPU fallback)")
        src = op.operand
        if getattr(src, "is_computed", False):
            compute_at(src, time)                                   # compute_at!(ψ, time), SpatialFilters.jl:354-356
        sd, dd = src.descriptor(), self.descriptor()
        ext = (C.c_int32 * 3)(*src.extent)
        lo = (C.c_int32 * 3)(*[w[0] for w in self.window])
        hi = (C.c_int32 * 3)(*[w[1] for w in self.window])
        dt = _lib.OCB_F64 if grid.dtype == torch.float64 else _lib.OCB_F32
        _lib.check(_lib.lib().ocb_filter_execute(dt, C.byref(sd), ext, self._passes, len(self._passes), C.byref(dd), lo, hi,
                                                 current_stream_ptr(grid.device)))
        self.fill_halo_regions()
        self.status_time = time
        return self


class _FilterOperator:
    """`BoxFilter(; dims, N, boundary)` / `GaussianFilter(; dims, σ, N, boundary)`: a callable ψ -> filter(ψ)."""

    def __init__(self, kind, **kw):
        self.kind, self.kw = kind, kw
        dims = tuplefy_dims(kw["dims"])
        validate_dims(dims)

    def __call__(self, psi):
        return FilterOperation(self.kind, psi, **self.kw)


def BoxFilter(psi=None, *, dims, N, boundary="shrink"):
    """box_filter.jl:91-99 (with ψ) and :201-206 (operator form)."""
    if psi is None:
        validate_N(N)
        return _FilterOperator("box", dims=dims, N=N, boundary=boundary)
    return FilterOperation("box", psi, dims=dims, N=N, boundary=boundary)


def GaussianFilter(psi=None, *, dims, sigma=None, N=None, boundary="shrink", σ=None):
    """gaussian_filter.jl:367-378 (with ψ) and :476-482 (operator form); `σ` is accepted as an alias of `sigma`."""
    sigma = σ if sigma is None else sigma
    if psi is None:
        if not isinstance(sigma, (int, float, np.floating)) or isinstance(sigma, bool) or not sigma > 0:
            raise ValueError(f"`σ` must be a positive number; got {sigma!r}")
        return _FilterOperator("gaussian", dims=dims, sigma=sigma, N=N, boundary=boundary)
    return FilterOperation("gaussian", psi, dims=dims, sigma=sigma, N=N, boundary=boundary)
