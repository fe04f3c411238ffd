"""AvailablePotentialEnergyEquation (src/AvailablePotentialEnergyEquation.jl): the local available potential energy,
the buoyancy displacement potential and the APE dissipation rate, built on the sorted reference state of
BackgroundPotentialEnergyEquation.  Scope table section 8(f), row 1 -- the consumers of the K3 sort.

Same constructors as the reference: `AvailablePotentialEnergy(model; method)` / `(model, z✶)`,
`BuoyancyDisplacementPotential(model; method)` / `(model, z✶)`, `AvailablePotentialEnergyDissipationRate(model; method)` /
`(model, z✶; upsilon)`; each returns a KernelFunctionOperation that `Field(op)` / `Integral(op)` accept."""
from __future__ import annotations

from . import _lib
from .fields import Field
from .operations import KernelFunctionOperation, SweepOperands
from .models import validate_location
from .BackgroundPotentialEnergyEquation import (BackgroundPotentialEnergy, SortedReferenceHeightField, SortedColumnField, reference_height,  # noqa: F401
                                                ThreeDimensionalSort, HeavisideIntegral, VerticalSort, ProfileLookup,
                                                validate_gravity_is_z_aligned)
from .TracerVarianceEquation import _tracer
from .KineticEnergyEquation import PotentialEnergyConversion as PotentialToKineticEnergyConversion  # noqa: F401  (w b)


def _zstar(name, model, zstar, method, default, column_ok=False):
    validate_gravity_is_z_aligned(name, model)
    if zstar is None:
        zstar = reference_height(model, method=method if method is not None else default)
    if not isinstance(zstar, (SortedReferenceHeightField, SortedColumnField)):
        raise TypeError(f"{name}(model, z✶) takes the field returned by reference_height")
    if isinstance(zstar, SortedColumnField):
        if not column_ok:                                 # validate_reference_height_grid
            raise ValueError(f"{name} needs a reference height on the model grid; a `VerticalSort` column lives on its own 1×1×N grid")
    elif zstar.grid is not model.grid:
        raise ValueError(f"{name}: the reference height was computed on another grid than the model's")
    return zstar


def AvailablePotentialEnergy(model, zstar=None, method=None, location=("c", "c", "c")):
    """e_a = Ψ(z) - Ψ(z✶) - b (z - z✶)  (local_ape_ccc, :42-45; constructors :99-111)."""
    validate_location(location, "AvailablePotentialEnergy")
    zstar = _zstar("AvailablePotentialEnergy", model, zstar, method, ThreeDimensionalSort(), column_ok=True)
    aux = {5: zstar.reference_potential, 6: zstar}
    flags = 0
    if isinstance(zstar, SortedColumnField):              # on a column the parcel's own height is a carried field (:43)
        aux[3] = zstar.sorted_height
        flags = _lib.OCB_REQ_CARRIED_HEIGHT
    return KernelFunctionOperation("APE", zstar.grid, SweepOperands(b=zstar.buoyancy, aux=aux), flags=flags,
                                   name="AvailablePotentialEnergy",
                                   computes="local available potential energy density  e_a = ∫ [b✶(z̃) - b] dz̃ from z✶ to z")


def BuoyancyDisplacementPotential(model, zstar=None, method=None, location=("c", "c", "c")):
    """Υ = z✶ - z  (upsilon_ccc, :126; constructors :180-192)."""
    validate_location(location, "BuoyancyDisplacementPotential")
    zstar = _zstar("BuoyancyDisplacementPotential", model, zstar, method, HeavisideIntegral())
    return KernelFunctionOperation("UPSILON", zstar.grid, SweepOperands(aux={6: zstar}), name="BuoyancyDisplacementPotential",
                                   computes="buoyancy displacement potential  Υ = z✶ - z")


def AvailablePotentialEnergyDissipationRate(model, zstar=None, method=None, upsilon=None, location=("c", "c", "c")):
    """ε_a = κ ∂ᵢb ∂ᵢΥ in the conservative form (1/V) Σ ℑ(-A δΥ qᵢ)  (ape_dissipation_rate_ccc, :202-215; constructors :282-301)."""
    validate_location(location, "AvailablePotentialEnergyDissipationRate")
    name = "AvailablePotentialEnergyDissipationRate"
    try:
        _tracer(model, "b")
    except Exception:
        raise ValueError(f"{name} needs buoyancy to be a tracer the closure diffuses (BuoyancyTracer with tracers=:b)")
    cl = model.closure
    if cl is None or (cl.kappa == 0.0 and not cl.kappa_fields):        # validate_closure_supplies_a_flux
        raise ValueError(f"{name} needs a closure that supplies a diffusive buoyancy flux (the model's has none)")
    zstar = _zstar(name, model, zstar, method, HeavisideIntegral())
    if upsilon is None:
        upsilon = Field(BuoyancyDisplacementPotential(model, zstar))
    if not isinstance(upsilon, Field) or upsilon.location != ("c", "c", "c"):
        raise TypeError("`upsilon` must be a Field at (Center, Center, Center)")
    return KernelFunctionOperation("APE_EPS", model.grid, SweepOperands(b=_tracer(model, "b"), aux={4: upsilon}, closure=model.closure),
                                   name=name, computes="available potential energy dissipation rate  ε_a = κ ∂ᵢb ∂ᵢΥ")


DissipationRate = AvailablePotentialEnergyDissipationRate

KineticEnergyConversion = PotentialToKineticEnergyConversion


def reference_buoyancy(zstar):
    """The buoyancy paired with a reference height: the model's own b on the model grid, the sorted profile b✶ on a
    `VerticalSort` column (BackgroundPotentialEnergyEquation.jl:270-296)."""
    return zstar.buoyancy


def PotentialEnergyDiffusiveVerticalBuoyancyFlux(*args, **kwargs):
    raise _lib.OcbError("PotentialEnergyDiffusiveVerticalBuoyancyFlux (src/PotentialEnergyEquation.jl) is outside the B200 hot path "
                        "(SURVEY.md section 8f, row 4): evaluate it with Oceanostics.jl")
