"""Sort-based background potential energy, restated in NumPy (test infrastructure only).

Follows src/BackgroundPotentialEnergyEquation.jl: rank_by_buoyancy! (:310-316), slot_centers!
(:326-331), slot_faces (:339-348), reference_potential_function (:359-368),
fill_reference_potential! (:383-397), assign_reference_height! for ThreeDimensionalSort (:413-426)
and HeavisideIntegral (:428-466), reference_height setup (:734-763), flat_grid_metric (:802-808).

Flat arrays are in the reference's column-major order (i fastest).  Sorting uses Julia's ``isless``
total order (-0.0 < +0.0, ties broken by original index as ``Base.sortperm!`` does); inputs are
assumed finite.
"""
from __future__ import annotations

import numpy as np

from .grid import OField


def order_key(b):
    """Order-preserving unsigned key of a float array (matches ``isless`` on finite values and ±0)."""
    if b.dtype == np.float64:
        u = b.view(np.uint64)
        sign = np.uint64(1) << np.uint64(63)
        return np.where(u & sign, ~u, u | sign)
    u = b.view(np.uint32)
    sign = np.uint32(1) << np.uint32(31)
    return np.where(u & sign, ~u, u | sign)


def flat_metrics(grid):
    i, j, k = grid.indices("ccc")
    shp = grid.extent_of("ccc")
    dV = np.broadcast_to(grid.V("ccc", i, j, k), shp).astype(grid.dtype)
    z = np.broadcast_to(grid.node(2, "c", k), shp).astype(grid.dtype)
    return dV.ravel(order="F").copy(), z.ravel(order="F").copy()


def reference_height(b: OField, method="ThreeDimensionalSort"):
    """-> dict(zstar (Nx,Ny,Nz), psi_delta (Nx,Ny,Nz), b_sorted, zstar_sorted, perm, A, z_bottom)."""
    grid = b.grid
    ft = grid.dtype.type
    shp = grid.extent_of("ccc")
    dV_flat, z_flat = flat_metrics(grid)
    if method != "HeavisideIntegral" and not all(grid.uniform):
        raise ValueError(f"`{method}` needs a grid with uniform cell volumes (regular spacing in every "
                         "direction), but this grid is stretched. Only `HeavisideIntegral()` supports stretched grids.")
    z_bottom = ft(grid.faces[2][1])
    A = ft(np.sum(dV_flat) / grid.L[2])

    work = b.interior.ravel(order="F").astype(grid.dtype).copy()
    perm = np.argsort(order_key(work), kind="stable")
    dV = dV_flat[perm]
    bs = work[perm]

    cum = np.cumsum(dV)
    if method == "ThreeDimensionalSort":
        zs_sorted = z_bottom + (cum - dV / 2) / A
    elif method == "HeavisideIntegral":
        below = cum - dV
        N = bs.size
        starts = np.ones(N, dtype=bool)
        starts[1:] = bs[1:] != bs[:-1]
        ends = np.ones(N, dtype=bool)
        ends[:-1] = starts[1:]
        denser = np.maximum.accumulate(np.where(starts, below, np.finfo(grid.dtype).min))
        no_lighter = np.minimum.accumulate(np.where(ends, cum, np.finfo(grid.dtype).max)[::-1])[::-1]
        zs_sorted = z_bottom + (denser + no_lighter) / (2 * A)
    else:
        raise NotImplementedError(method)
    zs_sorted = zs_sorted.astype(grid.dtype)

    zstar = np.empty_like(work)
    zstar[perm] = zs_sorted

    # Ψ (fill_reference_potential!)
    dz = dV / A
    faces = np.empty(bs.size + 1, dtype=grid.dtype)
    faces[0] = z_bottom
    faces[1:] = np.cumsum(dz) + z_bottom
    psi_face = np.empty_like(faces)
    psi_face[0] = 0
    psi_face[1:] = np.cumsum(bs * np.diff(faces))

    def psi(zeta):
        kk = np.clip(np.searchsorted(faces, zeta, side="right"), 1, bs.size)   # searchsortedlast, 1-based
        return psi_face[kk - 1] + bs[kk - 1] * (zeta - faces[kk - 1])

    psi_star = psi(zs_sorted)
    scat = np.empty_like(work)
    scat[perm] = psi_star
    psi_delta = psi(z_flat) - scat

    return dict(zstar=zstar.reshape(shp, order="F"), psi_delta=psi_delta.reshape(shp, order="F"),
                b_sorted=bs, zstar_sorted=zs_sorted, perm=perm, A=A, z_bottom=z_bottom,
                dV=dV_flat.reshape(shp, order="F"))


def vertical_sort(b: OField):
    """assign_reference_height!(::VerticalSort) (:584-597): the sorted column itself -- buoyancy densest first, the
    height each parcel came from, the column's own cell centres as z✶, and Ψδ = Ψ(source height) - Ψ(z✶) in sorted
    order (fill_reference_potential! with sorted_output = true, :383-397)."""
    r = reference_height(b, "ThreeDimensionalSort")
    _, z_flat = flat_metrics(b.grid)
    perm = r["perm"]
    return dict(zstar=r["zstar_sorted"], b_sorted=r["b_sorted"], sorted_height=z_flat[perm],
                psi_delta=r["psi_delta"].ravel(order="F")[perm], perm=perm)


def profile_lookup(b: OField, bstar, zprof):
    """assign_reference_height!(::ProfileLookup) (:550-582) for a supplied profile (b✶, z✶), with profile_slot_faces
    (:497-507) and lookup_profile's validation (:528-547).  -> dict(zstar, psi_delta) on the model grid."""
    grid = b.grid
    bstar, zprof = np.asarray(bstar, dtype=grid.dtype), np.asarray(zprof, dtype=grid.dtype)
    if bstar.size != zprof.size:
        raise ValueError(f"`ProfileLookup` was given a profile whose buoyancy and height have different lengths ({bstar.size} and {zprof.size}).")
    if np.any(np.diff(bstar) < 0):
        raise ValueError("`ProfileLookup` needs a reference profile ordered from the densest fluid up, but the buoyancy it was given is not non-decreasing.")
    if np.any(np.diff(zprof) < 0):
        raise ValueError("`ProfileLookup` needs the heights paired with `b✶` to rise with it, but the heights it was given are not non-decreasing.")
    M = bstar.size
    shp = grid.extent_of("ccc")
    dV_flat, z_flat = flat_metrics(grid)
    z_bottom = grid.dtype.type(grid.faces[2][1])
    A = np.sum(dV_flat) / grid.L[2]
    faces = np.empty(M + 1, dtype=grid.dtype)
    faces[0] = z_bottom
    faces[M] = z_bottom + np.sum(dV_flat) / A
    faces[1:M] = (zprof[:-1] + zprof[1:]) / 2
    work = b.interior.ravel(order="F").astype(grid.dtype)
    lo = np.clip(np.searchsorted(bstar, work, side="left") + 1, 1, M)            # searchsortedfirst, 1-based
    lol = np.maximum(lo - 1, 1)
    b_class = np.where(np.abs(bstar[lol - 1] - work) < np.abs(bstar[lo - 1] - work), bstar[lol - 1], bstar[lo - 1])
    run_first = np.searchsorted(bstar, b_class, side="left") + 1
    run_last = np.searchsorted(bstar, b_class, side="right")                      # searchsortedlast, 1-based
    zstar = (faces[run_first - 1] + faces[run_last]) / 2
    psi_face = np.empty(M + 1, dtype=np.float64)
    psi_face[0] = 0
    psi_face[1:] = np.cumsum(bstar.astype(np.float64) * np.diff(faces.astype(np.float64)))

    def psi(zeta):
        kk = np.clip(np.searchsorted(faces, zeta, side="right"), 1, M)
        return psi_face[kk - 1] + bstar[kk - 1] * (zeta - faces[kk - 1])

    psi_delta = psi(z_flat) - psi(zstar)
    return dict(zstar=zstar.astype(grid.dtype).reshape(shp, order="F"), psi_delta=psi_delta.astype(grid.dtype).reshape(shp, order="F"))
