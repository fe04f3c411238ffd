"""BackgroundPotentialEnergyEquation: the sorted reference state z✶ and e_b = -b z✶
(reference: src/BackgroundPotentialEnergyEquation.jl).

    z✶ = reference_height(model; method = ThreeDimensionalSort())     # a computed Field, recomputed by compute(z✶)
    E_b = BackgroundPotentialEnergy(model, z✶)                         # KernelFunctionOperation, -b z✶

`ThreeDimensionalSort` and `HeavisideIntegral` run on the GPU through `ocb_bpe_reference_height` (hand-written radix
sort + scan epilogues).  `VerticalSort` and `ProfileLookup` are the "next" rows of the scope table and raise.
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
    code = None


class ProfileLookup:
    code = None


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


def reference_height(model_or_b, method=None):
    """reference_height(model; method) / reference_height(b::Field; method) (:726-763): returns the computed z✶ field."""
    method = method if method is not None else ThreeDimensionalSort()
    if getattr(method, "code", None) is None:
        raise _lib.OcbError(f"{type(method).__name__} is not evaluated by the CUDA library yet (scope table 8f); use "
                            "ThreeDimensionalSort() or HeavisideIntegral()")
    if isinstance(model_or_b, Field):
        b = model_or_b
    else:
        model = model_or_b
        validate_gravity_is_z_aligned("reference_height", model)
        b = model.tracers.b
    if b.location != ("c", "c", "c"):
        raise ValueError("the buoyancy field must be located at (Center, Center, Center)")
    validate_grid_for_method(method, b.grid)
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
    if not isinstance(zstar, SortedReferenceHeightField):
        raise TypeError("BackgroundPotentialEnergy(model, z✶) takes the field returned by reference_height")
    return KernelFunctionOperation("BPE", zstar.grid, SweepOperands(b=zstar.buoyancy, aux={6: zstar}),
                                   name="BackgroundPotentialEnergy", computes="background potential energy per unit volume  e_b = -bz✶")
