"""TurbulentKineticEnergyEquation (reference: src/TurbulentKineticEnergyEquation.jl)."""
from __future__ import annotations

from . import _lib
from .fields import as_operand
from .models import validate_location
from .operations import KernelFunctionOperation, SweepOperands, perturbation_split
from . import KineticEnergyEquation as _KE

CCC = ("c", "c", "c")


def TurbulentKineticEnergy(model, u=None, v=None, w=None, U=None, V=None, W=None, location=CCC):
    """½ Σ (uᵢ - Uᵢ)², src/TurbulentKineticEnergyEquation.jl:54-60."""
    validate_location(location, "TurbulentKineticEnergy")
    if u is None:
        u, v, w = model.velocities.u, model.velocities.v, model.velocities.w
    return KernelFunctionOperation("TKE", model.grid, SweepOperands(u=u, v=v, w=w, U=as_operand(U), V=as_operand(V), W=as_operand(W)),
                                   name="TurbulentKineticEnergy", computes="turbulent kinetic energy  ½uᵢ′uᵢ′")


def TurbulentKineticEnergyIsotropicDissipationRate(model, U=None, V=None, W=None, **kw):
    """src/TurbulentKineticEnergyEquation.jl:89-97: the isotropic dissipation of (u-U, v-V, w-W)."""
    u, v, w = model.velocities.u, model.velocities.v, model.velocities.w
    return _KE.KineticEnergyIsotropicDissipationRate(u - as_operand(U), v - as_operand(V), w - as_operand(W), model.closure,
                                                     grid=model.grid, **kw)


def _shear(diag, name, computes):
    def ctor(*args, U=None, V=None, W=None, grid=None, location=CCC):
        validate_location(location, name)
        if len(args) == 1:          # (model; U, V, W): perturbations are the lazy u - U (:327-330)
            model = args[0]
            grid = model.grid
            up, vp, wp = model.velocities.u - as_operand(U), model.velocities.v - as_operand(V), model.velocities.w - as_operand(W)
        else:                       # (u′, v′, w′, U, V, W; grid)
            up, vp, wp, U, V, W = args
        ops, flags = {}, 0
        for slot, x in (("u", up), ("v", vp), ("w", wp)):
            f, mean = perturbation_split(x)
            ops[slot] = f
            if mean is not None:
                flags = _lib.OCB_REQ_PERTURBATION
                # the subtracted mean must be the mean whose gradient is taken: one operand slot serves both
                ops[slot.upper()] = as_operand(mean)
        for slot, m in (("U", U), ("V", V), ("W", W)):
            m = as_operand(m)
            if slot in ops and not (ops[slot] is m or (type(ops[slot]) is type(m) and getattr(m, "value", 0) == getattr(ops[slot], "value", 0)
                                                     and not hasattr(m, "data"))):
                raise _lib.OcbError("shear production with u′ = u - A and a different mean flow U is not evaluated in one sweep: "
                                    "materialise u′ with Field(u - A) first")
            ops[slot] = m
        grid = grid or getattr(ops["u"], "grid", None) or getattr(ops["v"], "grid", None)
        return KernelFunctionOperation(diag, grid, SweepOperands(**ops), flags=flags, name=name, computes=computes)
    ctor.__name__ = name
    return ctor


TurbulentKineticEnergyXShearProductionRate = _shear("SPX", "TurbulentKineticEnergyXShearProductionRate", "x shear production  -uᵢ′u′∂ₓUᵢ")
TurbulentKineticEnergyYShearProductionRate = _shear("SPY", "TurbulentKineticEnergyYShearProductionRate", "y shear production  -uᵢ′v′∂ᵧUᵢ")
TurbulentKineticEnergyZShearProductionRate = _shear("SPZ", "TurbulentKineticEnergyZShearProductionRate", "z shear production  -uᵢ′w′∂_zUᵢ")
TurbulentKineticEnergyShearProductionRate = _shear("SP", "TurbulentKineticEnergyShearProductionRate", "shear production  -uᵢ′uⱼ′∂ⱼUᵢ")
XShearProductionRate = TurbulentKineticEnergyXShearProductionRate
YShearProductionRate = TurbulentKineticEnergyYShearProductionRate
ZShearProductionRate = TurbulentKineticEnergyZShearProductionRate
ShearProductionRate = TurbulentKineticEnergyShearProductionRate
IsotropicDissipationRate = TurbulentKineticEnergyIsotropicDissipationRate
