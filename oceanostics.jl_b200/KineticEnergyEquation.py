"""KineticEnergyEquation: constructors of the KE budget terms (reference: src/KineticEnergyEquation.jl).
Same names, keyword arguments and error behaviour; each returns a KernelFunctionOperation that `Field`,
`compute`, `Integral`, `Average` and `FusedDiagnostics` accept."""
from __future__ import annotations

from . import _lib
from .fields import as_operand
from .models import perturbation_velocities, validate_location
from .operations import Closure, KernelFunctionOperation, SweepOperands, perturbation_split

CCC = ("c", "c", "c")


def _vel(velocities):
    return (velocities.u, velocities.v, velocities.w) if hasattr(velocities, "u") else tuple(velocities)


def KineticEnergy(model, u=None, v=None, w=None, location=CCC):
    """src/KineticEnergyEquation.jl:43-46, 68."""
    validate_location(location, "KineticEnergy")
    if u is None:
        u, v, w = _vel(model.velocities)
    return KernelFunctionOperation("KE", model.grid, SweepOperands(u=u, v=v, w=w), name="KineticEnergy",
                                   computes="kinetic energy  ½uᵢuᵢ")


def KineticEnergyAdvection(model, velocities=None, location=CCC):
    """src/KineticEnergyEquation.jl:184-187 (Centered second-order momentum advection)."""
    validate_location(location, "KineticEnergyAdvection")
    if getattr(model, "hydrostatic", False):
        raise TypeError("KineticEnergyAdvection is defined for a NonhydrostaticModel")
    from .models import advection_flags
    u, v, w = _vel(velocities if velocities is not None else model.velocities)
    return KernelFunctionOperation("ADV", model.grid, SweepOperands(u=u, v=v, w=w), flags=advection_flags(model, "KineticEnergyAdvection"),
                                   name="KineticEnergyAdvection",
                                   computes="kinetic energy advection  uᵢ∂ⱼ(uᵢuⱼ)")


def KineticEnergyPressureRedistribution(model, velocities=None, pressure=None, location=CCC):
    """src/KineticEnergyEquation.jl:353-358."""
    validate_location(location, "KineticEnergyPressureRedistribution")
    u, v, w = _vel(velocities if velocities is not None else model.velocities)
    if pressure is None:
        pressure = model.pressures.pNHS
        if getattr(model.pressures, "pHY", None) is not None:
            from .fields import Field
            pressure = Field(model.pressures.pHY + model.pressures.pNHS)   # sum(model.pressures)
    return KernelFunctionOperation("PR", model.grid, SweepOperands(u=u, v=v, w=w, p=pressure),
                                   name="KineticEnergyPressureRedistribution",
                                   computes="kinetic energy pressure redistribution  uᵢ∂ᵢp")


def PotentialEnergyConversion(model, velocities=None, tracers=None, location=CCC):
    """uᵢbᵢ, src/KineticEnergyEquation.jl:419-422 (BuoyancyTracer; the gravity unit vector may be tilted)."""
    validate_location(location, "PotentialEnergyConversion")
    u, v, w = _vel(velocities if velocities is not None else model.velocities)
    tracers = tracers if tracers is not None else model.tracers
    gu = model.buoyancy.gravity_unit_vector
    return KernelFunctionOperation("WB", model.grid, SweepOperands(u=u, v=v, w=w, b=tracers.b, ghat=tuple(-g for g in gu)),
                                   name="PotentialEnergyConversion", computes="potential to kinetic energy conversion  uᵢbᵢ")


def DissipationRate(model, U=None, V=None, W=None, location=CCC):
    """ε = ∂ⱼuᵢ·τᵢⱼ, src/KineticEnergyEquation.jl:482-493."""
    validate_location(location, "DissipationRate")
    ops = perturbation_velocities(model, U, V, W)
    pert = any(x is not None for x in (U, V, W))
    if not pert:        # no mean flow: the request does not read U, V, W, so it can share a sweep with siblings that carry means
        ops = {k: v for k, v in ops.items() if k in ("u", "v", "w")}
    return KernelFunctionOperation("EPS", model.grid, SweepOperands(closure=model.closure, **ops),
                                   flags=_lib.OCB_REQ_PERTURBATION if pert else 0, name="KineticEnergyDissipationRate",
                                   computes="kinetic energy dissipation rate  ε = ∂ⱼuᵢ·τᵢⱼ")


KineticEnergyDissipationRate = DissipationRate


def KineticEnergyIsotropicDissipationRate(*args, location=CCC, **kw):
    """ε = 2νSᵢⱼSᵢⱼ, src/KineticEnergyEquation.jl:540-550.  Call as (model) or (u, v, w, closure, grid=...)."""
    validate_location(location, "KineticEnergyIsotropicDissipationRate")
    if len(args) == 1:
        model = args[0]
        u, v, w = _vel(model.velocities)
        closure, grid = model.closure, model.grid
    else:
        u, v, w, closure = args[:4]
        grid = kw.get("grid") or getattr(perturbation_split(u)[0], "grid")
    if not isinstance(closure, Closure):
        raise ValueError(f"Cannot calculate dissipation rate for {closure}")
    flags, ops = 0, {}
    for name, x in (("u", u), ("v", v), ("w", w)):
        f, mean = perturbation_split(x)
        ops[name] = f
        if mean is not None:
            ops[name.upper()] = as_operand(mean)
            flags = _lib.OCB_REQ_PERTURBATION
    return KernelFunctionOperation("EPS_ISO", grid, SweepOperands(closure=closure, **ops), flags=flags,
                                   name="KineticEnergyIsotropicDissipationRate",
                                   computes="isotropic kinetic energy dissipation rate  ε = 2νSᵢⱼSᵢⱼ")


Advection = KineticEnergyAdvection
PressureRedistribution = KineticEnergyPressureRedistribution
BuoyancyProduction = PotentialEnergyConversion
IsotropicDissipationRate = KineticEnergyIsotropicDissipationRate


PotentialToKineticEnergyConversion = PotentialEnergyConversion


def _deferred(name):
    def raiser(*args, **kwargs):
        raise _lib.OcbError(f"{name} needs Oceananigans' full momentum tendency / stress-divergence / forcing kernels and is deferred "
                            "(SURVEY.md section 8f, row 2): evaluate it with Oceanostics.jl")
    raiser.__name__ = name
    return raiser


def KineticEnergyStress(model, location=CCC):
    """DIFF = uᵢ∂ⱼτᵢⱼ: each velocity times the divergence of its viscous fluxes, at cell centres
    (uᵢ∂ⱼ_τᵢⱼᶜᶜᶜ, src/KineticEnergyEquation.jl:191-202, constructor :232-246).  SURVEY.md section 8f, row 2."""
    validate_location(location, "KineticEnergyStress")
    u, v, w = _vel(model.velocities)
    return KernelFunctionOperation("KE_STRESS", model.grid, SweepOperands(u=u, v=v, w=w, closure=model.closure),
                                   name="KineticEnergyStress", computes="kinetic energy stress/diffusion  uᵢ∂ⱼτᵢⱼ")


def KineticEnergyForcing(model, location=CCC):
    """FORC = uᵢFᵤᵢ: each velocity times its forcing at the velocity point, at cell centres (uᵢFᵤᵢᶜᶜᶜ,
    src/KineticEnergyEquation.jl:251-260, constructor :289-297; NonhydrostaticModel only, as there).  The forcing functors
    are materialised by the host layer (`operations.ForcingField`).  SURVEY.md section 8f, row 2."""
    from .operations import forcing_operand
    validate_location(location, "KineticEnergyForcing")
    if model.hydrostatic:
        raise TypeError("KineticEnergyForcing is defined for NonhydrostaticModel only (src/KineticEnergyEquation.jl:289)")
    u, v, w = _vel(model.velocities)
    aux = {a: forcing_operand(getattr(model.forcing, n), f, model) for a, (n, f) in enumerate((("u", u), ("v", v), ("w", w)))}
    return KernelFunctionOperation("KE_FORCING", model.grid, SweepOperands(u=u, v=v, w=w, aux=aux),
                                   name="KineticEnergyForcing", computes="kinetic energy forcing  uᵢFᵤᵢ")


def KineticEnergyTendency(model, location=CCC):
    """KET = uᵢGᵢ, the kinetic-energy tendency excluding the nonhydrostatic pressure contribution (uᵢGᵢᶜᶜᶜ,
    src/KineticEnergyEquation.jl:74-95, constructor :120-142): each velocity times Oceananigans' NonhydrostaticModel velocity
    tendency -- advection (Centered second order), Coriolis (FPlane / ConstantCartesianCoriolis), buoyancy, hydrostatic
    pressure anomaly, viscous-flux divergence, forcing -- for a model without background fields, immersed boundaries or
    Stokes drift (anything else is refused).  SURVEY.md section 8f, row 2."""
    from .models import get_coriolis_frequency_components
    from .operations import forcing_operand
    validate_location(location, "KineticEnergyTendency")
    if model.hydrostatic:
        raise TypeError("KineticEnergyTendency is defined for NonhydrostaticModel only (src/KineticEnergyEquation.jl:120)")
    from .models import advection_flags
    flags0 = advection_flags(model, "KineticEnergyTendency")
    for name in ("stokes_drift", "background_fields", "immersed_boundary"):
        if getattr(model, name, None) is not None:
            raise _lib.OcbError(f"KineticEnergyTendency with model.{name} is not evaluated by the CUDA library")
    u, v, w = _vel(model.velocities)
    kw = dict(u=u, v=v, w=w, closure=model.closure, f=tuple(get_coriolis_frequency_components(model.coriolis)))
    if model.buoyancy is not None:
        kw["b"] = model.tracers.b
        kw["ghat"] = tuple(-g for g in model.buoyancy.gravity_unit_vector)
    flags = flags0
    pHY = getattr(model.pressures, "pHY", None)
    if pHY is not None:
        kw["p"] = pHY
        flags |= _lib.OCB_REQ_HYDROSTATIC_SPLIT
    kw["aux"] = {a: forcing_operand(getattr(model.forcing, n), f_, model) for a, (n, f_) in enumerate((("u", u), ("v", v), ("w", w)))}
    return KernelFunctionOperation("KE_TENDENCY", model.grid, SweepOperands(**kw), flags=flags, name="KineticEnergyTendency",
                                   computes="kinetic energy tendency  uᵢGᵢ (excl. nonhydrostatic pressure)")


Stress = KineticEnergyStress
Forcing = KineticEnergyForcing
Tendency = KineticEnergyTendency
