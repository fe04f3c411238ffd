"""Oceananigans staggered-grid operators, restated for the NumPy oracle (test infrastructure only).

[OCN-mem: Oceananigans v0.10x Operators/{difference,interpolation,derivative,
spacings_and_areas_and_volumes}_operators.jl].  Naming: the Julia name with the
superscripts spelled out, e.g.  δxᶜᵃᵃ -> dx_caa, ℑxyᶜᶜᵃ -> ixy_cca, ∂zᶠᶜᶠ -> ddz_fcf,
Axᶜᶜᶜ -> Ax_ccc.  Every operator has the two Julia method families in one Python
function: ``op(i, j, k, grid, field)`` and ``op(i, j, k, grid, f, *args)`` with
``f(i, j, k, grid, *args)`` a kernel function.

  δxᶜᵃᵃ(i, u) = u[i+1] - u[i]        δxᶠᵃᵃ(i, c) = c[i] - c[i-1]
  ℑxᶜᵃᵃ(i, u) = (u[i] + u[i+1]) / 2  ℑxᶠᵃᵃ(i, c) = (c[i-1] + c[i]) / 2
  ℑxyᶜᶜᵃ f    = ℑyᵃᶜᵃ(ℑxᶜᵃᵃ f)     (outer operator = the later direction)
  ∂xᶜᶜᶜ u     = δxᶜᵃᵃ u / Δxᶜᶜᶜ     (a true division, at the result location)
"""
from __future__ import annotations

import itertools

_AX = {"x": 0, "y": 1, "z": 2}


def _ev(f, i, j, k, grid, *args):
    if callable(f):
        return f(i, j, k, grid, *args)
    return f[i, j, k]


def _sh(d, i, j, k, s):
    if d == 0:
        return i + s, j, k
    if d == 1:
        return i, j + s, k
    return i, j, k + s


def _make_delta(d, l):
    lo, hi = (0, 1) if l == "c" else (-1, 0)

    def op(i, j, k, grid, f, *args):
        return _ev(f, *_sh(d, i, j, k, hi), grid, *args) - _ev(f, *_sh(d, i, j, k, lo), grid, *args)
    return op


def _make_interp(d, l):
    lo, hi = (0, 1) if l == "c" else (-1, 0)

    def op(i, j, k, grid, f, *args):
        return (_ev(f, *_sh(d, i, j, k, lo), grid, *args) + _ev(f, *_sh(d, i, j, k, hi), grid, *args)) / 2
    return op


def _name1(dname, l):
    s = ["a", "a", "a"]
    s[_AX[dname]] = l
    return "".join(s)


_g = globals()
for _dn in "xyz":
    for _l in "cf":
        _g[f"d{_dn}_{_name1(_dn, _l)}"] = _make_delta(_AX[_dn], _l)
        _g[f"i{_dn}_{_name1(_dn, _l)}"] = _make_interp(_AX[_dn], _l)


def _make_interp2(d1, l1, d2, l2):
    inner = _g[f"i{d1}_{_name1(d1, l1)}"]
    outer = _g[f"i{d2}_{_name1(d2, l2)}"]

    def op(i, j, k, grid, f, *args):
        return outer(i, j, k, grid, inner, f, *args)
    return op


for (_d1, _d2) in (("x", "y"), ("x", "z"), ("y", "z")):
    for _l1, _l2 in itertools.product("cf", "cf"):
        s = ["a", "a", "a"]
        s[_AX[_d1]] = _l1
        s[_AX[_d2]] = _l2
        _g[f"i{_d1}{_d2}_{''.join(s)}"] = _make_interp2(_d1, _l1, _d2, _l2)


def _make_metric(kind, loc):
    def op(i, j, k, grid, *unused):
        return getattr(grid, kind)(loc, i, j, k)
    return op


for _loc in itertools.product("cf", "cf", "cf"):
    _ls = "".join(_loc)
    for _kind in ("Dx", "Dy", "Dz", "Ax", "Ay", "Az", "V"):
        _g[f"{_kind}_{_ls}"] = _make_metric(_kind, _loc)


def _make_deriv(dn, loc):
    d = _AX[dn]
    delta = _g[f"d{dn}_{_name1(dn, loc[d])}"]
    kind = "D" + dn

    def op(i, j, k, grid, f, *args):
        return delta(i, j, k, grid, f, *args) / getattr(grid, kind)(loc, i, j, k)
    return op


for _loc in itertools.product("cf", "cf", "cf"):
    for _dn in "xyz":
        _g[f"dd{_dn}_{''.join(_loc)}"] = _make_deriv(_dn, _loc)
