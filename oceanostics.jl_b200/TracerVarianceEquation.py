"""TracerVarianceEquation (reference: src/TracerVarianceEquation.jl)."""
from __future__ import annotations

from . import _lib
from .fields import as_operand
from .models import validate_location
from .operations import KernelFunctionOperation, SweepOperands, perturbation_split

CCC = ("c", "c", "c")


def _tracer(model, tracer_name):
    name = str(tracer_name).lstrip(":")
    if not hasattr(model.tracers, name):
        raise KeyError(f"model has no tracer named {name}")
    return getattr(model.tracers, name)


def TracerVarianceDissipationRate(model, tracer_name, tracer=None, location=CCC):
    """χ = 2 ∂ⱼc·κ∂ⱼc, src/TracerVarianceEquation.jl:210-226 (`tracer` may be a lazy c - C)."""
    validate_location(location, "TracerVarianceDissipationRate")
    c = tracer if tracer is not None else _tracer(model, tracer_name)
    c, mean = perturbation_split(c)
    ops = dict(b=c)
    flags = 0
    if mean is not None:
        ops["B"] = as_operand(mean)
        flags = _lib.OCB_REQ_PERTURBATION
    return KernelFunctionOperation("CHI", model.grid, SweepOperands(closure=model.closure, **ops), flags=flags,
                                   name="TracerVarianceDissipationRate", computes="tracer variance dissipation rate  χ = 2∂ⱼc·Fⱼ")


def TracerVarianceDiffusion(model, tracer_name, location=CCC):
    """2c ∇·q, src/TracerVarianceEquation.jl:129-136."""
    validate_location(location, "TracerVarianceDiffusion")
    return KernelFunctionOperation("CDIFF", model.grid, SweepOperands(closure=model.closure, b=_tracer(model, tracer_name)),
                                   name="TracerVarianceDiffusion", computes="tracer variance diffusion  2c∂ⱼFⱼ")


DissipationRate = TracerVarianceDissipationRate
Diffusion = TracerVarianceDiffusion



def TracerVarianceTendency(model, tracer_name, location=CCC):
    """TEND = 2c ∂ₜc with Oceananigans' NonhydrostaticModel tracer tendency (c∂ₜcᶜᶜᶜ, src/TracerVarianceEquation.jl:18-39,
    constructor :66-88):  ∂ₜc = -div_Uc - ∇·q + F  for a model with Centered(order = 2 | 4) advection and without background
    fields, biogeochemistry or immersed boundaries (this model container has none); the tracer forcing is materialised by
    the host layer (`operations.ForcingField`).  SURVEY.md section 8f, row 2."""
    from .operations import forcing_operand
    validate_location(location, "TracerVarianceTendency")
    if model.hydrostatic:
        raise TypeError("TracerVarianceTendency is defined for NonhydrostaticModel only (src/TracerVarianceEquation.jl:66)")
    from .models import advection_flags
    flags = advection_flags(model, "TracerVarianceTendency")
    c = _tracer(model, tracer_name)
    name = str(tracer_name).lstrip(":")
    u, v, w = model.velocities.u, model.velocities.v, model.velocities.w
    forcing = forcing_operand(getattr(model.forcing, name, None), c, model)
    return KernelFunctionOperation("C_TENDENCY", model.grid, SweepOperands(u=u, v=v, w=w, b=c, closure=model.closure, aux={0: forcing}),
                                   flags=flags, name="TracerVarianceTendency", computes="tracer variance tendency  2c ∂ₜc")


Tendency = TracerVarianceTendency
