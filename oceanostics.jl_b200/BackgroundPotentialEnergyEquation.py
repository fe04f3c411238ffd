"""BackgroundPotentialEnergyEquation: the sorted reference state z✶ and e_b = -b z✶
(reference: src/BackgroundPotentialEnergyEquation.jl).

    z✶ = reference_height(model; method = ThreeDimensionalSort())     # a computed Field, recomputed by compute(z✶)
    E_b = BackgroundPotentialEnergy(model, z✶)                         # KernelFunctionOperation, -b z✶

`ThreeDimensionalSort` and `HeavisideIntegral` run on the GPU through `ocb_bpe_reference_height` (hand-written radix
sort + scan epilogues).  `VerticalSort` reads the same sort out in sorted order onto a 1 x 1 x N column
(`ocb_gather_sorted`); `ProfileLookup` matches every cell to a sorted profile by binary search
(`ocb_bpe_profile_lookup`).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from .fields import CenterField, Field, current_stream_ptr
from .models import validate_location
from .operations import KernelFunctionOperation, SweepOperands, compute_at


class ThreeDimensionalSort:
    code = _lib.OCB_BPE_THREE_DIMENSIONAL_SORT

    def __repr__(self):
        return "ThreeDimensionalSort()"


class HeavisideIntegral:
    code = _lib.OCB_BPE_HEAVISIDE_INTEGRAL

    def __repr__(self):
        return "HeavisideIntegral()"


class VerticalSort:
    """The sorted column itself, on a 1 x 1 x N grid (:152-157, 584-597)."""
    code = None

    def __repr__(self):
        return "VerticalSort()"


class ProfileLookup:
    """ProfileLookup() sorts the field's own buoyancy; ProfileLookup(z✶_column) borrows a VerticalSort column;
    ProfileLookup(b✶, z✶) takes any paired buoyancy and height ordered from the densest fluid up (:224-229)."""
    code = None

    def __init__(self, *profile):
        if len(profile) == 0:
            self.profile = None
        elif len(profile) == 1:
            self.profile = profile[0]
        elif len(profile) == 2:
            self.profile = (profile[0], profile[1])
        else:
            raise TypeError("ProfileLookup takes no argument, a VerticalSort reference height, or a (b✶, z✶) pair")

    def __repr__(self):
        return "ProfileLookup()" if self.profile is None else "ProfileLookup (external profile)"


def validate_grid_for_method(method, grid):
    """src/BackgroundPotentialEnergyEquation.jl:770-788: the stacking methods need uniform cell volumes."""
    if isinstance(method, HeavisideIntegral):
        return
    if not all(grid.uniform):
        raise ValueError(f"`{method!r}` needs a grid with uniform cell volumes (regular spacing in every direction), but this "
                         "grid is stretched. Only `HeavisideIntegral()` supports stretched grids.")


def validate_gravity_is_z_aligned(name, model):
    g = getattr(model.buoyancy, "gravity_unit_vector", (0.0, 0.0, -1.0))
    if g[0] != 0 or g[1] != 0:
        raise ValueError(f"{name} requires gravity aligned with the z axis; got gravity_unit_vector = {g}")


class SortedReferenceHeightField(Field):
    """`Field` whose operand is the SortedReferenceState of a buoyancy field (:71-81); `compute` re-sorts (:610-619)."""
    is_computed = True

    def __new__(cls, *a, **k):
        return object.__new__(cls)

    def __init__(self, b: Field, method):
        grid = b.grid
        Field.__init__(self, ("c", "c", "c"), grid)
        self.buoyancy, self.method = b, method
        self.reference_potential = CenterField(grid)      # Ψδ = Ψ(z) - Ψ(z✶), used by the APE diagnostics
        self.reference_potential.boundary["impenetrable"] = False
        self.boundary["impenetrable"] = False
        n = grid.N[0] * grid.N[1] * grid.N[2]
        self.permutation = torch.zeros(n, dtype=torch.int64, device=grid.device)     # sortperm (1-based), :757
        self.sorted_buoyancy = torch.zeros(n, dtype=grid.dtype, device=grid.device)
        self.group = None

    def compute(self, time=None):
        grid = self.grid
        if grid.device.type != "cuda":
            raise _lib.OcbError("the sorted reference state is computed by the CUDA library only (no CPU fallback)")
        if getattr(self.buoyancy, "is_computed", False):
            compute_at(self.buoyancy, time)
        gd, bd, zd, pd = grid.descriptor(), self.buoyancy.descriptor(), self.descriptor(), self.reference_potential.descriptor()
        _lib.check(_lib.lib().ocb_bpe_reference_height(C.byref(gd), C.byref(bd), self.method.code, C.byref(zd), C.byref(pd),
                                                       self.permutation.data_ptr(), self.sorted_buoyancy.data_ptr(),
                                                       current_stream_ptr(grid.device)))
        self.fill_halo_regions()
        self.reference_potential.fill_halo_regions()
        self.status_time = time
        return self


def _dtype_code(grid):
    return _lib.OCB_F64 if grid.dtype == torch.float64 else _lib.OCB_F32


def _gather_sorted(field, perm, out):
    """out[r] = field[perm[r]] (sorted order) through the library's gather kernel."""
    d = field.descriptor()
    _lib.check(_lib.lib().ocb_gather_sorted(_dtype_code(field.grid), C.byref(d), perm.data_ptr(), out.data_ptr(), out.numel(),
                                            current_stream_ptr(field.grid.device)))
    return out


class SortedColumnField(Field):
    """z✶ of `VerticalSort`: the sorted cells stacked on a 1 x 1 x N column that keeps the model grid's topology and
    spans its horizontal area (build_sorting_method, :826-847).  Carries `reference_buoyancy` (b✶, densest first),
    `sorted_height` (where each parcel came from) and `reference_potential` (Ψδ in sorted order)."""
    is_computed = True

    def __new__(cls, *a, **k):
        return object.__new__(cls)

    def __init__(self, b: Field, method):
        from .grids import RectilinearGrid, Flat
        g = b.grid
        self._ranked = SortedReferenceHeightField(b, ThreeDimensionalSort())
        n = g.N[0] * g.N[1] * g.N[2]
        topo = g.user_topology
        size = tuple(s for s, t in zip((1, 1, n), topo) if t != Flat)
        zf = g.nodes(2, "f", [1, g.N[2] + 1])
        kw = {}
        if topo[0] != Flat:
            kw["x"] = (0.0, g.L[0])
        if topo[1] != Flat:
            kw["y"] = (0.0, g.L[1])
        column = RectilinearGrid(size, z=(float(zf[0]), float(zf[0]) + g.L[2]), halo=tuple(h for h, t in zip(g.H, topo) if t != Flat),
                                 topology=topo, dtype=g.dtype, device=g.device, **kw)
        Field.__init__(self, ("c", "c", "c"), column)
        self.method, self.source_buoyancy, self.model_grid = method, b, g
        self.reference_buoyancy, self.sorted_height = CenterField(column), CenterField(column)
        self.reference_potential = CenterField(column)
        for f in (self, self.reference_buoyancy, self.sorted_height, self.reference_potential):
            f.boundary["impenetrable"] = False
        self.buoyancy = self.reference_buoyancy           # reference_buoyancy(z✶, ::VerticalSort, b) = the column's b✶ (:287)
        method.reference_buoyancy, method.sorted_height = self.reference_buoyancy, self.sorted_height
        self._heights = CenterField(g)                    # Zᶜᶜᶜ of the model grid
        zc = torch.from_numpy(g.nodes(2, "c", list(range(1, g.N[2] + 1))).astype(g.np_dtype)).to(g.device)
        self._heights.interior.copy_(zc.reshape(-1, 1, 1).expand_as(self._heights.interior))
        self._tmp = torch.empty(n, dtype=g.dtype, device=g.device)
        self.group = None

    @property
    def permutation(self):
        return self._ranked.permutation

    def compute(self, time=None):
        r = self._ranked
        r.compute(time)
        perm = r.permutation
        n = self._tmp.numel()
        for src, dst in ((r, self), (self._heights, self.sorted_height), (r.reference_potential, self.reference_potential)):
            _gather_sorted(src, perm, self._tmp)
            dst.interior.copy_(self._tmp.reshape(n, 1, 1))
        self.reference_buoyancy.interior.copy_(r.sorted_buoyancy.reshape(n, 1, 1))
        for f in (self, self.reference_buoyancy, self.sorted_height, self.reference_potential):
            f.fill_halo_regions()
        self.status_time = time
        return self


class ProfileLookupHeightField(SortedReferenceHeightField):
    """z✶ of `ProfileLookup`: every cell takes the mid-height of the run of profile slots of its buoyancy class
    (assign_reference_height!, :550-582)."""

    def __init__(self, b: Field, method):
        SortedReferenceHeightField.__init__(self, b, method)
        p = method.profile
        self._ranked = SortedReferenceHeightField(b, ThreeDimensionalSort()) if p is None else None
        if p is not None and not isinstance(p, (SortedColumnField, tuple)):
            raise TypeError("`ProfileLookup` takes either a reference height built with `VerticalSort()`, or a pair of buoyancies and "
                            f"heights given as `ProfileLookup(b✶, z✶)`. It was handed a {type(p).__name__}, which is neither.")
        n = b.grid.N[0] * b.grid.N[1] * b.grid.N[2]
        self._zprof = torch.empty(n, dtype=b.grid.dtype, device=b.grid.device) if p is None else None

    def _profile(self, time):
        g, p = self.grid, self.method.profile
        if p is None:                                     # lookup_profile(::ProfileLookup{Nothing}), :517-526
            self._ranked.compute(time)
            return self._ranked.sorted_buoyancy, _gather_sorted(self._ranked, self._ranked.permutation, self._zprof)
        if isinstance(p, SortedColumnField):              # refresh_profile!: the borrowed column tracks the flow (:600-608)
            compute_at(p, time)
            return p.reference_buoyancy.interior.reshape(-1).contiguous(), p.interior.reshape(-1).contiguous()
        to_dev = lambda a: (a.interior.reshape(-1) if isinstance(a, Field) else torch.as_tensor(a)).to(device=g.device, dtype=g.dtype).contiguous()
        return to_dev(p[0]), to_dev(p[1])

    def compute(self, time=None):
        g = self.grid
        if g.device.type != "cuda":
            raise _lib.OcbError("the sorted reference state is computed by the CUDA library only (no CPU fallback)")
        if getattr(self.buoyancy, "is_computed", False):
            compute_at(self.buoyancy, time)
        bs, zp = self._profile(time)
        if bs.numel() != zp.numel():                      # lookup_profile's validation, :528-547
            raise ValueError(f"`ProfileLookup` was given a profile whose buoyancy and height have different lengths ({bs.numel()} and {zp.numel()}).")
        if bs.numel() > 1 and not bool(torch.all(bs[1:] >= bs[:-1])):
            raise ValueError("`ProfileLookup` needs a reference profile ordered from the densest fluid up, but the buoyancy it was given "
                             "is not non-decreasing.")
        if zp.numel() > 1 and not bool(torch.all(zp[1:] >= zp[:-1])):
            raise ValueError("`ProfileLookup` needs the heights paired with `b✶` to rise with it, but the heights it was given are not "
                             "non-decreasing. A reference profile runs from the densest fluid at the bottom to the lightest at the top.")
        self._keep = (bs, zp)
        gd, bd, zd, pd = g.descriptor(), self.buoyancy.descriptor(), self.descriptor(), self.reference_potential.descriptor()
        _lib.check(_lib.lib().ocb_bpe_profile_lookup(C.byref(gd), C.byref(bd), bs.data_ptr(), zp.data_ptr(), bs.numel(), C.byref(zd), C.byref(pd),
                                                     current_stream_ptr(g.device)))
        self.fill_halo_regions()
        self.reference_potential.fill_halo_regions()
        self.status_time = time
        return self


def reference_height(model_or_b, method=None):
    """reference_height(model; method) / reference_height(b::Field; method) (:726-763): returns the computed z✶ field."""
    method = method if method is not None else ThreeDimensionalSort()
    if not isinstance(method, (ThreeDimensionalSort, HeavisideIntegral, VerticalSort, ProfileLookup)):
        raise TypeError(f"unknown reference-height method {method!r}")
    if isinstance(model_or_b, Field):
        b = model_or_b
    else:
        model = model_or_b
        validate_gravity_is_z_aligned("reference_height", model)
        b = model.tracers.b
    if b.location != ("c", "c", "c"):
        raise ValueError("the buoyancy field must be located at (Center, Center, Center)")
    validate_grid_for_method(method, b.grid)
    if isinstance(method, VerticalSort):
        z = SortedColumnField(b, method)
    elif isinstance(method, ProfileLookup):
        z = ProfileLookupHeightField(b, method)
    else:
        z = SortedReferenceHeightField(b, method)
    if b.grid.device.type == "cuda":
        z.compute()
    return z


def BackgroundPotentialEnergy(model, zstar=None, method=None, location=("c", "c", "c")):
    """e_b = -b z✶ (:899-907)."""
    validate_location(location, "BackgroundPotentialEnergy")
    validate_gravity_is_z_aligned("BackgroundPotentialEnergy", model)
    if zstar is None:
        zstar = reference_height(model, method=method)
    if not isinstance(zstar, (SortedReferenceHeightField, SortedColumnField)):
        raise TypeError("BackgroundPotentialEnergy(model, z✶) takes the field returned by reference_height")
    return KernelFunctionOperation("BPE", zstar.grid, SweepOperands(b=zstar.buoyancy, aux={6: zstar}),
                                   name="BackgroundPotentialEnergy", computes="background potential energy per unit volume  e_b = -bz✶")
