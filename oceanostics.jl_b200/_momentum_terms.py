"""Shared builder of the pass-through modules UMomentumEquation / VMomentumEquation / WMomentumEquation (reference:
src/UMomentumEquation.jl, src/VMomentumEquation.jl, src/WMomentumEquation.jl -- thin constructors over Oceananigans' own
tendency kernels).  Each term is one request of the library's OCB_DIAG_U_TERM / V_TERM / W_TERM evaluator with the term code in
the request flags; the model must be a NonhydrostaticModel without background fields, immersed boundaries or Stokes drift
(what the library restates of Oceananigans' `u_velocity_tendency`), with Centered second-order advection."""
from __future__ import annotations

from . import _lib
from .models import get_coriolis_frequency_components, validate_location
from .operations import KernelFunctionOperation, SweepOperands, forcing_operand

_COMP = {"u": ("U_TERM", ("f", "c", "c"), 0, "x"), "v": ("V_TERM", ("c", "f", "c"), 1, "y"), "w": ("W_TERM", ("c", "c", "f"), 2, "z")}


def _flags(term):
    return term << _lib.OCB_REQ_TERM_SHIFT


def _check_model(model, what):
    if getattr(model, "hydrostatic", False):
        raise _lib.OcbError(f"{what}: the HydrostaticFreeSurfaceModel forms are not evaluated by the CUDA library; use Oceanostics.jl")
    for name in ("stokes_drift", "background_fields", "immersed_boundary"):
        if getattr(model, name, None) is not None:
            raise _lib.OcbError(f"{what} with model.{name} is not evaluated by the CUDA library")


def _check_advection(model):
    """-> the request flag of the model's advection scheme (Centered(order = 2 | 4)); refuses anything else."""
    from .models import advection_flags
    return advection_flags(model, "advection term")


def make_module(comp):
    """-> dict of constructors for the `comp`-momentum equation (comp in 'u', 'v', 'w')."""
    diag, loc, axis, xyz = _COMP[comp]
    C = comp.upper()

    def vel(model):
        v = model.velocities
        return dict(u=v.u, v=v.v, w=v.w)

    def op(model, term, name, computes, **operands):
        return KernelFunctionOperation(diag, model.grid, SweepOperands(**vel(model), **operands), flags=_flags(term),
                                       name=f"{C}{name}", computes=computes)

    def Advection(model, location=loc):
        f"""ADV = ∂ⱼ(uⱼ{comp}): Oceananigans' div_𝐯{comp} (src/{C}MomentumEquation.jl `Advection`)."""
        validate_location(location, "Advection", loc)
        _check_model(model, "Advection")
        o = op(model, _lib.OCB_TERM_ADVECTION, "Advection", f"advection of {comp}-momentum  ∂ⱼ(uⱼ{comp})")
        o.flags |= _check_advection(model)
        return o

    def BuoyancyAcceleration(model, location=loc):
        validate_location(location, "BuoyancyAcceleration", loc)
        if model.buoyancy is None:
            raise ValueError("BuoyancyAcceleration needs a model with buoyancy")
        ghat = tuple(-g for g in model.buoyancy.gravity_unit_vector)
        return op(model, _lib.OCB_TERM_BUOYANCY, "BuoyancyAcceleration", f"buoyancy acceleration in {xyz}  ĝ b", b=model.tracers.b, ghat=ghat)

    def CoriolisAcceleration(model, location=loc):
        validate_location(location, "CoriolisAcceleration", loc)
        f = tuple(get_coriolis_frequency_components(model.coriolis))
        return op(model, _lib.OCB_TERM_CORIOLIS, "CoriolisAcceleration", f"Coriolis acceleration in {xyz}  (f × u)", f=f)

    def PressureGradient(model, hydrostatic_pressure=None, location=loc):
        validate_location(location, "PressureGradient", loc)
        p = hydrostatic_pressure if hydrostatic_pressure is not None else getattr(model.pressures, "pHY", None)
        kw = {} if p is None else {"p": p}                          # hydrostatic_pressure_gradient of `nothing` is zero
        return op(model, _lib.OCB_TERM_PRESSURE, "PressureGradient", f"hydrostatic pressure gradient in {xyz}", **kw)

    def ViscousDissipation(model, location=loc):
        validate_location(location, "ViscousDissipation", loc)
        return op(model, _lib.OCB_TERM_DIFFUSION, "ViscousDissipation", f"viscous flux divergence  ∂ⱼτ_{axis + 1}ⱼ", closure=model.closure)

    def TotalViscousDissipation(model, location=loc):
        _check_model(model, "TotalViscousDissipation")              # without an immersed boundary the total is the interior term
        return ViscousDissipation(model, location=location)

    def Forcing(model, location=loc):
        validate_location(location, "Forcing", loc)
        forced = getattr(model.velocities, comp)
        return op(model, _lib.OCB_TERM_FORCING, "Forcing", f"forcing of {comp}", aux={axis: forcing_operand(getattr(model.forcing, comp), forced, model)})

    def Tendency(model, location=loc):
        validate_location(location, "Tendency", loc)
        _check_model(model, "Tendency")
        adv_flag = _check_advection(model)
        v = model.velocities
        kw = dict(closure=model.closure, f=tuple(get_coriolis_frequency_components(model.coriolis)))
        if model.buoyancy is not None:
            kw["b"] = model.tracers.b
            kw["ghat"] = tuple(-g for g in model.buoyancy.gravity_unit_vector)
        flags = _flags(_lib.OCB_TERM_TENDENCY) | adv_flag
        pHY = getattr(model.pressures, "pHY", None)
        if pHY is not None:
            kw["p"] = pHY
            flags |= _lib.OCB_REQ_HYDROSTATIC_SPLIT
        kw["aux"] = {a: forcing_operand(getattr(model.forcing, n), getattr(v, n), model) for a, n in enumerate("uvw")}
        return KernelFunctionOperation(diag, model.grid, SweepOperands(**vel(model), **kw), flags=flags, name=f"{C}Tendency",
                                       computes=f"total tendency of {comp}")

    def _unsupported(name):
        def raiser(*a, **k):
            raise _lib.OcbError(f"{C}MomentumEquation.{name} (free surface / immersed boundary / Stokes drift terms) is not evaluated "
                                "by the CUDA library: evaluate it with Oceanostics.jl")
        raiser.__name__ = name
        return raiser

    out = dict(Advection=Advection, BuoyancyAcceleration=BuoyancyAcceleration, CoriolisAcceleration=CoriolisAcceleration,
               PressureGradient=PressureGradient, ViscousDissipation=ViscousDissipation, TotalViscousDissipation=TotalViscousDissipation,
               Forcing=Forcing, Tendency=Tendency)
    for name in ("BarotropicPressureGradient", "ImmersedViscousDissipation", "StokesShear", "StokesTendency"):
        out[name] = _unsupported(name)
    out.update({f"{C}{k}": v for k, v in list(out.items())})       # UAdvection, UBuoyancyAcceleration, ... (the reference's aliases)
    return out
