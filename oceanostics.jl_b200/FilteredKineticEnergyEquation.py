"""FilteredKineticEnergyEquation: kinetic energy budget terms of the low-pass filtered flow
(reference: src/FilteredKineticEnergyEquation.jl).  Host-side graph builders: they materialise the filtered
velocities / fluxes with the staged filters (K2) and contract them in one sweep kernel (K1)."""
from __future__ import annotations

from types import SimpleNamespace

from . import _lib
from .fields import Field
from . import FlowDiagnostics
from .operations import KernelFunctionOperation, SweepOperands
from .SpatialFilters import GaussianFilter


def _filter_from_kwargs(filter, kw):
    if filter is not None:
        return filter
    sigma = kw.get("sigma", kw.get("σ"))
    return GaussianFilter(dims=kw.get("dims", (1, 2, 3)), sigma=sigma, boundary=kw.get("boundary", "shrink"), N=kw.get("N"))


def filtered_velocities(filter, dims, u, v, w):
    """:29-34: velocity component d is filtered iff d ∈ dims; each filtered velocity is a materialised Field."""
    ub = Field(filter(u)) if 1 in dims else u
    vb = Field(filter(v)) if 2 in dims else v
    wb = Field(filter(w)) if 3 in dims else w
    return ub, vb, wb


def FilteredKineticEnergy(model, filter=None, **kw):
    """Kˡ = ½ ūᵢūᵢ (:98-104)."""
    filter = _filter_from_kwargs(filter, kw)
    u, v, w = model.velocities.u, model.velocities.v, model.velocities.w
    ub, vb, wb = filtered_velocities(filter, (1, 2, 3), u, v, w)
    return KernelFunctionOperation("KE", model.grid, SweepOperands(u=ub, v=vb, w=wb), name="FilteredKineticEnergy",
                                   computes="kinetic energy of the filtered flow  Kˡ = ½ūᵢūᵢ")


_TENSOR_KEYS = (("τ₁₁", (1, 1), 0), ("τ₂₂", (2, 2), 1), ("τ₃₃", (3, 3), 2), ("τ₁₂", (1, 2), 3), ("τ₁₃", (1, 3), 4), ("τ₂₃", (2, 3), 5))


def _filtered_momentum_fluxes(filter, grid, u, v, w, dims, collocate_diagonals):
    """filter(uⁱuʲ) for the kept components, as materialised Fields keyed by aux slot (:42-48)."""
    full = FlowDiagnostics.StressTensor(grid, u, v, w, dims=dims, collocate_diagonals=collocate_diagonals)
    out = {}
    for key, _, slot in _TENSOR_KEYS:
        if hasattr(full, key):
            out[slot] = Field(filter(Field(getattr(full, key))))
    return out


def subfilter_stress_tensor(model, filter=None, dims=(1, 2, 3), collocate_diagonals=False, **kw):
    """τⁱʲ = filter(uⁱuʲ) - ūⁱūʲ component by component, each at its StressTensor location (:42-48, :118-160): a namespace
    of computed Fields keyed τ₁₁ ... τ₂₃ (the kept components for `dims`)."""
    from .operations import PointwiseField
    FlowDiagnostics.validate_dims(dims)
    filter = _filter_from_kwargs(filter, dict(kw, dims=kw.get("filter_dims", dims)))
    grid = model.grid
    u, v, w = model.velocities.u, model.velocities.v, model.velocities.w
    ub, vb, wb = filtered_velocities(filter, dims, u, v, w)
    full = FlowDiagnostics.StressTensor(grid, u, v, w, dims=dims, collocate_diagonals=collocate_diagonals)
    filt = FlowDiagnostics.StressTensor(grid, ub, vb, wb, dims=dims, collocate_diagonals=collocate_diagonals)
    out = {}
    for key, _, _slot in _TENSOR_KEYS:
        if hasattr(full, key):
            out[key] = PointwiseField("-", Field(filter(Field(getattr(full, key)))), Field(getattr(filt, key)))
    return SimpleNamespace(**out)


def KineticEnergyCrossScaleFlux(model, filter=None, dims=(1, 2, 3), collocate_diagonals=False, **kw):
    """Πₖ = -τⁱʲ S̄ⁱʲ contracted at cell centres (:176-188, :246-261); τⁱʲ = filter(uⁱuʲ) - ūⁱūʲ."""
    FlowDiagnostics.validate_dims(dims)
    filter = _filter_from_kwargs(filter, dict(kw, dims=kw.get("filter_dims", dims)))
    grid = model.grid
    u, v, w = model.velocities.u, model.velocities.v, model.velocities.w
    ub, vb, wb = filtered_velocities(filter, dims, u, v, w)
    aux = _filtered_momentum_fluxes(filter, grid, u, v, w, dims, collocate_diagonals)
    flags = sum(f for d, f in ((1, _lib.OCB_REQ_TENSOR_DIM1), (2, _lib.OCB_REQ_TENSOR_DIM2), (3, _lib.OCB_REQ_TENSOR_DIM3)) if d in dims)
    if collocate_diagonals:
        flags |= _lib.OCB_REQ_COLLOCATED
    return KernelFunctionOperation("PI", grid, SweepOperands(u=ub, v=vb, w=wb, aux=aux), flags=flags, name="KineticEnergyCrossScaleFlux",
                                   computes="cross-scale kinetic energy flux  Πₖ = -τⁱʲS̄ⁱʲ")


def _filtered_viscous_fluxes(model, filter):
    """τ̄ᵢⱼ = filter(τᵢⱼ(u)): the full-flow viscous fluxes, each filtered at its staggered location (:377-387).
    viscous_flux_uy = viscous_flux_vx etc., so six distinct fields serve the nine terms."""
    u, v, w = model.velocities.u, model.velocities.v, model.velocities.w
    ops = SweepOperands(u=u, v=v, w=w, closure=model.closure)
    return {slot: Field(filter(Field(KernelFunctionOperation(diag, model.grid, ops, name=diag))))
            for slot, diag in enumerate(("VF11", "VF22", "VF33", "VF12", "VF13", "VF23"))}


def FilteredKineticEnergyDissipationRate(model, filter=None, **kw):
    """εˡ = ∂ⱼūᵢ·τ̄ᵢⱼ (:275-299, :366-389)."""
    filter = _filter_from_kwargs(filter, kw)
    u, v, w = model.velocities.u, model.velocities.v, model.velocities.w
    ub, vb, wb = Field(filter(u)), Field(filter(v)), Field(filter(w))
    aux = _filtered_viscous_fluxes(model, filter)
    return KernelFunctionOperation("FEPS", model.grid, SweepOperands(u=ub, v=vb, w=wb, aux=aux), name="FilteredKineticEnergyDissipationRate",
                                   computes="filtered-flow kinetic energy dissipation rate  εˡ = ∂ⱼūᵢ·τ̄ᵢⱼ")


DissipationRate = FilteredKineticEnergyDissipationRate
CrossScaleFlux = KineticEnergyCrossScaleFlux
