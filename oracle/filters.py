"""Separable box / Gaussian spatial filters, restated in NumPy (test infrastructure only).

Follows src/SpatialFilters/SpatialFilters.jl (policies :26-149, extent rule :152-162, policy
resolution :166-202, validation :234-271, staged compute :347-378), box_filter.jl (:19-47, :91-99)
and gaussian_filter.jl (uniform :44-81, stretched :193-262, weights :312, kernel choice :337-356,
constructor :367-378, widths :488-505).  One 1-D pass accumulates taps in the same order
(Δ = -w .. w) and with the same ``s += w*val; w_sum += w*cnt`` recurrences as the Julia kernels.
"""
from __future__ import annotations

import math

import numpy as np

from .grid import OField


# --- validation (SpatialFilters.jl:234-271) ----------------------------------------------------------
def tuplefy_dims(dims):
    return (int(dims),) if isinstance(dims, (int, np.integer)) and not isinstance(dims, bool) else dims


def validate_dims(dims):
    if not isinstance(dims, tuple):
        raise ValueError(f"`dims` must be a tuple of integers; got {type(dims).__name__}")
    ok = (len(dims) > 0 and all(isinstance(d, (int, np.integer)) and d in (1, 2, 3) for d in dims)
          and len(set(dims)) == len(dims))
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


def resolve_filter_policies(field, dims, boundary):
    """-> (sorted_dims, policies) with policies[i] = (kind, [left, right], extent)."""
    validate_dims(dims)
    grid = field.grid
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
        spec = specs[dims.index(d)]
        base = ("periodic",) if grid.topology[d - 1] == "P" else parse_boundary_spec(spec)
        policies.append(base + (field.ext[d - 1],))
    return sorted_dims, tuple(policies)


def validate_periodic_widths(grid, sorted_dims, policies, widths):
    for d, pol, w in zip(sorted_dims, policies, widths):
        if pol[0] == "periodic":
            Nd = grid.N[d - 1]
            N = 2 * w + 1
            if N > 2 * Nd + 1:
                raise ValueError(f"for the periodic direction d={d}, `N` ({N}) exceeds the maximum allowed "
                                 f"2*Nd_grid+1 = {2 * Nd + 1}")


# --- one 1-D pass ------------------------------------------------------------------------------------
def _taps(policy, n, width):
    """For each offset t=-w..w: (index array into 0..n-1, value-mask, count) following the readers
    at SpatialFilters.jl:101-124."""
    kind = policy[0]
    base = np.arange(1, n + 1)
    for t in range(-width, width + 1):
        m = base + t
        if kind == "periodic":
            mr = m + n * (m < 1) - n * (m > n)
            yield t, m, mr - 1, None, np.ones(n, dtype=np.int64)
        elif kind == "edge":
            yield t, m, np.clip(m, 1, n) - 1, None, np.ones(n, dtype=np.int64)
        elif kind == "constant":
            yield t, m, np.clip(m, 1, n) - 1, ("const", m < 1, m > n), np.ones(n, dtype=np.int64)
        elif kind == "shrink":
            inb = (m >= 1) & (m <= n)
            yield t, m, np.clip(m, 1, n) - 1, ("zero", ~inb), inb.astype(np.int64)
        else:
            raise ValueError(kind)


def _fetch(a, d, idx, special, policy):
    val = np.take(a, idx, axis=d)
    if special is None:
        return val
    shp = [1, 1, 1]
    shp[d] = -1
    if special[0] == "zero":
        return np.where(special[1].reshape(shp), a.dtype.type(0), val)
    lo, hi = special[1].reshape(shp), special[2].reshape(shp)
    return np.where(lo, a.dtype.type(policy[1]), np.where(hi, a.dtype.type(policy[2]), val))


def box_pass(a, d, width, policy):
    """BoxFilterKernel{D} (box_filter.jl:19-47): s / n with integer n."""
    n = a.shape[d]
    s = np.zeros_like(a)
    cnt_tot = np.zeros(n, dtype=np.int64)
    for t, m, idx, special, cnt in _taps(policy, n, width):
        s = s + _fetch(a, d, idx, special, policy)
        cnt_tot += cnt
    shp = [1, 1, 1]
    shp[d] = -1
    return (s / cnt_tot.reshape(shp).astype(a.dtype)).astype(a.dtype)


def gaussian_weights(width, sigma_cells, dtype=np.float64):
    """gaussian_filter.jl:312 (weights are computed in the grid's FT)."""
    ft = np.dtype(dtype).type
    sig = ft(sigma_cells)
    return [ft(np.exp(-ft((idx - width - 1) ** 2) / (ft(2) * sig * sig)))
            for idx in range(1, 2 * width + 2)]


def gaussian_pass_uniform(a, d, width, policy, weights):
    """GaussianFilterKernel{D} (gaussian_filter.jl:44-81)."""
    n = a.shape[d]
    ft = a.dtype.type
    s = np.zeros_like(a)
    wsum = np.zeros(n, dtype=a.dtype)
    for (t, m, idx, special, cnt), w in zip(_taps(policy, n, width), weights):
        s = s + ft(w) * _fetch(a, d, idx, special, policy)
        wsum = wsum + ft(w) * cnt.astype(a.dtype)
    shp = [1, 1, 1]
    shp[d] = -1
    with np.errstate(invalid="ignore", divide="ignore"):
        return (s / wsum.reshape(shp)).astype(a.dtype)


def gaussian_pass_stretched(a, d, width, policy, sigma, nodes, spacings, period):
    """StretchedGaussianFilterKernel{D} (gaussian_filter.jl:193-262).  ``nodes`` / ``spacings`` are the
    interior node coordinates / widths along d at the operand's location (length n)."""
    n = a.shape[d]
    ft = a.dtype.type
    sigma = ft(sigma)
    x0 = nodes.astype(a.dtype)
    s = np.zeros_like(a)
    wsum = np.zeros(n, dtype=a.dtype)
    shp = [1, 1, 1]
    shp[d] = -1
    for t, m, idx, special, cnt in _taps(policy, n, width):
        if policy[0] == "periodic":
            x = nodes[idx] - ft(period) * (m < 1) + ft(period) * (m > n)
        else:
            x = nodes[idx]
        w = (spacings[idx] * np.exp(-(x - x0) ** 2 / (2 * sigma ** 2))).astype(a.dtype)
        s = s + w.reshape(shp) * _fetch(a, d, idx, special, policy)
        wsum = wsum + w * cnt.astype(a.dtype)
    with np.errstate(invalid="ignore", divide="ignore"):
        return (s / wsum.reshape(shp)).astype(a.dtype)


# --- drivers (staged evaluation, SpatialFilters.jl:347-378) ---------------------------------------------
def box_filter(field: OField, dims, N, boundary="shrink"):
    dims = tuplefy_dims(dims)
    validate_N(N)
    sorted_dims, policies = resolve_filter_policies(field, dims, boundary)
    width = (N - 1) // 2
    validate_periodic_widths(field.grid, sorted_dims, policies, [width] * len(sorted_dims))
    a = field.interior.copy()
    for d, pol in zip(sorted_dims, policies):
        a = box_pass(a, d - 1, width, pol)
    return a


def infer_width(sigma, grid, d):
    return int(math.ceil(2 * sigma / float(np.min(_interior_spacings(grid, d)))))


def _interior_spacings(grid, d):
    return grid.dc[d - 1][np.arange(1, grid.N[d - 1] + 1)]


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


def gaussian_filter(field: OField, dims, sigma, N=None, boundary="shrink"):
    dims = tuplefy_dims(dims)
    if not isinstance(sigma, (int, float, np.floating)) or isinstance(sigma, bool) or not sigma > 0:
        raise ValueError(f"`σ` must be a positive number; got {sigma!r}")
    grid = field.grid
    sorted_dims, policies = resolve_filter_policies(field, dims, boundary)
    widths = resolve_gaussian_widths(N, sigma, grid, dims, sorted_dims)
    validate_periodic_widths(grid, sorted_dims, policies, widths)
    ft = grid.dtype.type
    a = field.interior.copy()
    for d, pol, w in zip(sorted_dims, policies, widths):
        sp = _interior_spacings(grid, d)
        if sp.min() == sp.max():   # direction_is_uniform, gaussian_filter.jl:337-340
            sigma_cells = ft(sigma) / ft(sp.min())
            a = gaussian_pass_uniform(a, d - 1, w, pol, gaussian_weights(w, sigma_cells, grid.dtype))
        else:
            l = field.loc[d - 1]
            idx = np.arange(1, field.ext[d - 1] + 1)
            nodes = grid.node(d - 1, l, idx)
            spac = grid.spacing(d - 1, l, idx)
            a = gaussian_pass_stretched(a, d - 1, w, pol, sigma, nodes, spac, grid.L[d - 1])
    return a
