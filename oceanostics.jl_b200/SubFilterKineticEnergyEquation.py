"""SubFilterKineticEnergyEquation (reference: src/SubFilterKineticEnergyEquation.jl): Kˢ = filter(K) - Kˡ and
εˢ = filter(ε) - εˡ."""
from __future__ import annotations

from .fields import Field
from . import KineticEnergyEquation as _KE
from .FilteredKineticEnergyEquation import _filter_from_kwargs, _filtered_viscous_fluxes, filtered_velocities
from .operations import KernelFunctionOperation, SweepOperands


def SubFilterKineticEnergy(model, filter=None, **kw):
    """:76-85."""
    filter = _filter_from_kwargs(filter, kw)
    u, v, w = model.velocities.u, model.velocities.v, model.velocities.w
    kbar = Field(filter(Field(_KE.KineticEnergy(model, u, v, w))))
    ub, vb, wb = filtered_velocities(filter, (1, 2, 3), u, v, w)
    return KernelFunctionOperation("SFKE", model.grid, SweepOperands(u=ub, v=vb, w=wb, aux={6: kbar}), name="SubFilterKineticEnergy",
                                   computes="sub-filter kinetic energy  Kˢ = ½τⁱⁱ")


def SubFilterKineticEnergyDissipationRate(model, filter=None, **kw):
    """:143-151."""
    filter = _filter_from_kwargs(filter, kw)
    u, v, w = model.velocities.u, model.velocities.v, model.velocities.w
    ebar = Field(filter(Field(_KE.DissipationRate(model))))
    ub, vb, wb = Field(filter(u)), Field(filter(v)), Field(filter(w))
    aux = dict(_filtered_viscous_fluxes(model, filter))
    aux[6] = ebar
    return KernelFunctionOperation("SFEPS", model.grid, SweepOperands(u=ub, v=vb, w=wb, aux=aux), name="SubFilterKineticEnergyDissipationRate",
                                   computes="sub-filter kinetic energy dissipation rate  εˢ = filter(ε) - εˡ")


DissipationRate = SubFilterKineticEnergyDissipationRate


from .FilteredKineticEnergyEquation import KineticEnergyCrossScaleFlux  # noqa: E402,F401  (re-exported as in the reference)
