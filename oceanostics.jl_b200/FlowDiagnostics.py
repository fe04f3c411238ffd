"""FlowDiagnostics (reference: src/FlowDiagnostics.jl): Ri, Ro, Ertel PV family, strain / vorticity / stress tensors."""
from __future__ import annotations

from types import SimpleNamespace

from . import _lib
from .fields import normalize_location
from .models import get_coriolis_frequency_components
from .operations import KernelFunctionOperation, SweepOperands

CCC, CCF, FFF = ("c", "c", "c"), ("c", "c", "f"), ("f", "f", "f")


def _validate_location(loc, name, valid):
    if normalize_location(loc) != tuple(valid):
        raise ValueError(f"{name} only supports location = {valid} for now.")


def validate_dims(dims):
    """src/FlowDiagnostics.jl validate_dims: a non-empty tuple of distinct integers from (1, 2, 3)."""
    ok = (isinstance(dims, tuple) and len(dims) > 0 and all(isinstance(d, int) and d in (1, 2, 3) for d in dims)
          and len(set(dims)) == len(dims))
    if not ok:
        raise ValueError(f"`dims` must be a non-empty tuple of distinct integers drawn from (1, 2, 3); got {dims}")


def RichardsonNumber(model, u=None, v=None, w=None, b=None, true_vertical_direction=None, loc=CCF):
    """src/FlowDiagnostics.jl:98-120: Ri = ∂b/∂ẑ / |∂uₕ/∂ẑ|² with ẑ the anti-gravity direction."""
    _validate_location(loc, "RichardsonNumber", CCF)
    if u is None:
        u, v, w, b = model.velocities.u, model.velocities.v, model.velocities.w, model.tracers.b
    if true_vertical_direction is None:
        true_vertical_direction = tuple(-g for g in model.buoyancy.gravity_unit_vector)
    return KernelFunctionOperation("RI", model.grid, SweepOperands(u=u, v=v, w=w, b=b, dir=tuple(true_vertical_direction)),
                                   name="RichardsonNumber", computes="Richardson number  Ri = N²/S²")


def RossbyNumber(model, u=None, v=None, w=None, coriolis=None, loc=FFF, dWdy_bg=0, dVdz_bg=0, dUdz_bg=0, dWdx_bg=0,
                 dUdy_bg=0, dVdx_bg=0):
    """src/FlowDiagnostics.jl:168-196."""
    _validate_location(loc, "RossbyNumber", FFF)
    if u is None:
        u, v, w = model.velocities.u, model.velocities.v, model.velocities.w
    fx, fy, fz = get_coriolis_frequency_components(coriolis if coriolis is not None else model.coriolis)
    return KernelFunctionOperation("RO", model.grid, SweepOperands(u=u, v=v, w=w, f=(fx, fy, fz),
                                                                    ro_bg=(dWdy_bg, dVdz_bg, dUdz_bg, dWdx_bg, dUdy_bg, dVdx_bg)),
                                   name="RossbyNumber", computes="Rossby number  Ro = ω·f/|f|²")


def ErtelPotentialVorticity(model, u=None, v=None, w=None, tracer=None, coriolis=None, tracer_name="b", thermal_wind=False, loc=FFF):
    """src/FlowDiagnostics.jl:333-356."""
    _validate_location(loc, "ErtelPotentialVorticity", FFF)
    if u is None:
        u, v, w = model.velocities.u, model.velocities.v, model.velocities.w
    if tracer is None:
        tracer = getattr(model.tracers, str(tracer_name).lstrip(":"))
    fx, fy, fz = get_coriolis_frequency_components(coriolis if coriolis is not None else model.coriolis)
    if thermal_wind:
        return KernelFunctionOperation("TWPV", model.grid, SweepOperands(u=u, v=v, b=tracer, f=(fx, fy, fz)),
                                       name="ErtelPotentialVorticity", computes="potential vorticity (thermal wind)")
    return KernelFunctionOperation("EPV", model.grid, SweepOperands(u=u, v=v, w=w, b=tracer, f=(fx, fy, fz)),
                                   name="ErtelPotentialVorticity", computes="Ertel potential vorticity  q = (f + ω)·∇b")


def DirectionalErtelPotentialVorticity(model, direction, u=None, v=None, w=None, tracer=None, coriolis=None, tracer_name="b", loc=FFF):
    """src/FlowDiagnostics.jl:410-425."""
    _validate_location(loc, "DirectionalErtelPotentialVorticity", FFF)
    if u is None:
        u, v, w = model.velocities.u, model.velocities.v, model.velocities.w
    if tracer is None:
        tracer = getattr(model.tracers, str(tracer_name).lstrip(":"))
    fx, fy, fz = get_coriolis_frequency_components(coriolis if coriolis is not None else model.coriolis)
    f_dir = fx * direction[0] + fy * direction[1] + fz * direction[2]
    return KernelFunctionOperation("DEPV", model.grid, SweepOperands(u=u, v=v, w=w, b=tracer, dir=tuple(direction), f_dir=f_dir),
                                   name="DirectionalErtelPotentialVorticity", computes="directional Ertel potential vorticity")


def _uvw(model_or_grid, u, v, w):
    if u is None:
        m = model_or_grid
        return m.grid, m.velocities.u, m.velocities.v, m.velocities.w
    return model_or_grid, u, v, w


def StrainRateTensorModulus(model, loc=CCC):
    """src/FlowDiagnostics.jl:475-478."""
    _validate_location(loc, "StrainRateTensorModulus", CCC)
    g, u, v, w = _uvw(model, None, None, None)
    return KernelFunctionOperation("SMOD", g, SweepOperands(u=u, v=v, w=w), name="StrainRateTensorModulus",
                                   computes="strain rate tensor modulus  √(SᵢⱼSᵢⱼ)")


def VorticityTensorModulus(model, loc=CCC):
    """src/FlowDiagnostics.jl:621-624."""
    _validate_location(loc, "VorticityTensorModulus", CCC)
    g, u, v, w = _uvw(model, None, None, None)
    return KernelFunctionOperation("OMOD", g, SweepOperands(u=u, v=v, w=w), name="VorticityTensorModulus",
                                   computes="vorticity tensor modulus  √(ΩᵢⱼΩᵢⱼ)")


def QVelocityGradientTensorInvariant(model, loc=CCC):
    """src/FlowDiagnostics.jl:949-954."""
    _validate_location(loc, "QVelocityGradientTensorInvariant", CCC)
    g, u, v, w = _uvw(model, None, None, None)
    return KernelFunctionOperation("Q", g, SweepOperands(u=u, v=v, w=w), name="QVelocityGradientTensorInvariant",
                                   computes="Q invariant  ½(ΩᵢⱼΩᵢⱼ - SᵢⱼSᵢⱼ)")


def _tensor(prefix, pairs, grid, u, v, w, dims):
    validate_dims(dims)
    out = {}
    sub = {1: "₁", 2: "₂", 3: "₃"}
    sym = {"S": "S", "O": "Ω", "T": "τ"}[prefix[0]]
    for (i, j), diag in pairs.items():
        if i in dims and j in dims:
            out[f"{sym}{sub[i]}{sub[j]}"] = KernelFunctionOperation(diag, grid, SweepOperands(u=u, v=v, w=w), name=f"{sym}{sub[i]}{sub[j]}")
    return SimpleNamespace(**out)


def StrainRateTensor(model_or_grid, u=None, v=None, w=None, dims=(1, 2, 3)):
    """src/FlowDiagnostics.jl:555-571: NamedTuple (here a namespace) of the kept components."""
    g, u, v, w = _uvw(model_or_grid, u, v, w)
    return _tensor("S", {(1, 1): "S11", (2, 2): "S22", (3, 3): "S33", (1, 2): "S12", (1, 3): "S13", (2, 3): "S23"}, g, u, v, w, dims)


def VorticityTensor(model_or_grid, u=None, v=None, w=None, dims=(1, 2, 3)):
    """src/FlowDiagnostics.jl:692-705."""
    g, u, v, w = _uvw(model_or_grid, u, v, w)
    return _tensor("O", {(1, 2): "O12", (1, 3): "O13", (2, 3): "O23"}, g, u, v, w, dims)


def StressTensor(model_or_grid, u=None, v=None, w=None, dims=(1, 2, 3), collocate_diagonals=False):
    """uᵢuⱼ at natural locations, src/FlowDiagnostics.jl:826-851."""
    g, u, v, w = _uvw(model_or_grid, u, v, w)
    c = "C" if collocate_diagonals else ""
    return _tensor("T", {(1, 1): "T11" + c, (2, 2): "T22" + c, (3, 3): "T33" + c, (1, 2): "T12", (1, 3): "T13", (2, 3): "T23"},
                   g, u, v, w, dims)


def ThermalWindPotentialVorticity(model, **kw):
    """potential_vorticity_in_thermal_wind_fff (src/FlowDiagnostics.jl:200-214, 259): ErtelPotentialVorticity's thermal-wind form."""
    return ErtelPotentialVorticity(model, thermal_wind=True, **kw)


def to_center(psi):
    """`@at (Center, Center, Center) ψ` (src/FlowDiagnostics.jl:858) as a computed Field."""
    from .operations import PointwiseField
    from .fields import Field as _Field
    psi = psi if isinstance(psi, _Field) else _Field(psi)
    return psi if psi.location == ("c", "c", "c") and not getattr(psi, "is_reduced", False) else PointwiseField("@", psi, location=("c", "c", "c"))


def subfilter_covariance(a, b, filter, loc=("c", "c", "c"), filtered_a=None):
    """τ(a, b) = filter(a·b) - filter(a)·filter(b), co-located at `loc` (src/FlowDiagnostics.jl:901-910).  Every filtered
    piece is a materialised Field on the staged-filter path; the result is a computed Field that refreshes the whole
    graph on `compute`.  `filtered_a` shares a pre-filtered first factor between several covariances."""
    from .operations import PointwiseField
    from .fields import Field as _Field
    loc = normalize_location(loc)

    def at_loc(x):
        x = x if isinstance(x, _Field) else _Field(x)
        return x if x.location == loc else PointwiseField("@", x, location=loc)
    a_loc, b_loc = at_loc(a), at_loc(b)
    a_bar = _Field(filter(a_loc)) if filtered_a is None else filtered_a
    filtered_product = _Field(filter(PointwiseField("*", a_loc, b_loc, location=loc)))
    return PointwiseField("-", filtered_product, PointwiseField("*", a_bar, _Field(filter(b_loc)), location=loc), location=loc)


def _outside_the_hot_path(name, where):
    def raiser(*args, **kwargs):
        raise _lib.OcbError(f"{name} ({where}) is outside the B200 hot path (SURVEY.md section 8f / 2): evaluate it with Oceanostics.jl")
    raiser.__name__ = name
    return raiser


MixedLayerDepth = _outside_the_hot_path("MixedLayerDepth", "src/FlowDiagnostics.jl")
BuoyancyAnomalyCriterion = _outside_the_hot_path("BuoyancyAnomalyCriterion", "src/FlowDiagnostics.jl")
DensityAnomalyCriterion = _outside_the_hot_path("DensityAnomalyCriterion", "src/FlowDiagnostics.jl")
BottomCellValue = _outside_the_hot_path("BottomCellValue", "src/FlowDiagnostics.jl")
