"""TracerEquation: the terms of a tracer's evolution equation one by one (reference: src/TracerEquation.jl:70-327) -- Advection
(div_Uc), Diffusion (∇_dot_qᶜ), TotalDiffusion, X/Y/ZDiffusiveFlux, Forcing.  SURVEY.md section 8f row 4 (pass-through modules)."""
from __future__ import annotations

from . import _lib
from ._momentum_terms import _check_advection, _check_model, _flags
from .models import validate_location
from .operations import KernelFunctionOperation, SweepOperands, forcing_operand

CCC = ("c", "c", "c")


def _tracer(model, tracer_name):
    if not hasattr(model.tracers, tracer_name):
        raise ValueError(f"{tracer_name!r} is not a tracer of the model")
    return getattr(model.tracers, tracer_name)


def _vel(model):
    v = model.velocities
    return dict(u=v.u, v=v.v, w=v.w)


def Advection(model, tracer_name, location=CCC):
    """∂ⱼ(uⱼ c) = div_Uc, Centered(order = 2 | 4) (src/TracerEquation.jl:70-78)."""
    validate_location(location, "Advection")
    _check_model(model, "TracerEquation.Advection")
    return KernelFunctionOperation("C_TERM", model.grid, SweepOperands(b=_tracer(model, tracer_name), **_vel(model)),
                                   flags=_flags(_lib.OCB_TERM_ADVECTION) | _check_advection(model), name="TracerAdvection", computes="tracer advection  ∂ⱼ(uⱼ c)")


def Diffusion(model, tracer_name, location=CCC):
    """∂ⱼ qᶜⱼ = ∇_dot_qᶜ of the model's closure (src/TracerEquation.jl:112-120)."""
    validate_location(location, "Diffusion")
    return KernelFunctionOperation("C_TERM", model.grid, SweepOperands(b=_tracer(model, tracer_name), closure=model.closure, **_vel(model)),
                                   flags=_flags(_lib.OCB_TERM_DIFFUSION), name="TracerDiffusion", computes="tracer diffusion  ∂ⱼqᶜⱼ")


def TotalDiffusion(model, tracer_name, location=CCC):
    """Interior + immersed diffusion (src/TracerEquation.jl:186-194); without an immersed boundary, the interior term."""
    _check_model(model, "TracerEquation.TotalDiffusion")
    return Diffusion(model, tracer_name, location=location)


def _flux(diag, loc, axis):
    def ctor(model, tracer_name, location=loc):
        validate_location(location, f"{axis.upper()}DiffusiveFlux", loc)
        return KernelFunctionOperation(diag, model.grid, SweepOperands(b=_tracer(model, tracer_name), closure=model.closure),
                                       name=f"Tracer{axis.upper()}DiffusiveFlux", computes=f"diffusive tracer flux in {axis}  qᶜ")
    ctor.__name__ = f"{axis.upper()}DiffusiveFlux"
    ctor.__doc__ = f"diffusive_flux_{axis} of the model's closure (src/TracerEquation.jl:221-294)."
    return ctor


XDiffusiveFlux = _flux("QX", ("f", "c", "c"), "x")
YDiffusiveFlux = _flux("QY", ("c", "f", "c"), "y")
ZDiffusiveFlux = _flux("QZ", ("c", "c", "f"), "z")


def Forcing(model, tracer_name, location=CCC):
    """The tracer's forcing at the cell centres (src/TracerEquation.jl:322-330)."""
    validate_location(location, "Forcing")
    c = _tracer(model, tracer_name)
    return KernelFunctionOperation("C_TERM", model.grid, SweepOperands(b=c, aux={0: forcing_operand(getattr(model.forcing, tracer_name), c, model)},
                                                                      **_vel(model)),
                                   flags=_flags(_lib.OCB_TERM_FORCING), name="TracerForcing", computes="tracer forcing")


def ImmersedDiffusion(*a, **k):
    raise _lib.OcbError("TracerEquation.ImmersedDiffusion (immersed boundaries) is not evaluated by the CUDA library: evaluate it with Oceanostics.jl")


TracerAdvection, TracerDiffusion, TracerTotalDiffusion, TracerImmersedDiffusion = Advection, Diffusion, TotalDiffusion, ImmersedDiffusion
TracerXDiffusiveFlux, TracerYDiffusiveFlux, TracerZDiffusiveFlux, TracerForcing = XDiffusiveFlux, YDiffusiveFlux, ZDiffusiveFlux, Forcing
