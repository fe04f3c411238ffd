"""A minimal model container: the fields, closure and parameters the diagnostics constructors read from an
Oceananigans `NonhydrostaticModel` / `HydrostaticFreeSurfaceModel` (`model.velocities`, `model.tracers`,
`model.pressures`, `model.closure`, `model.closure_fields`, `model.coriolis`, `model.buoyancy`,
`model.advection`, `model.clock`).  It holds state only -- time stepping stays in Oceananigans."""
from __future__ import annotations

from types import SimpleNamespace

from .fields import CenterField, XFaceField, YFaceField, ZFaceField, ZeroField
from .operations import Closure


class FPlane:
    def __init__(self, f):
        self.f = float(f)


class ConstantCartesianCoriolis:
    def __init__(self, fx=0.0, fy=0.0, fz=0.0):
        self.fx, self.fy, self.fz = float(fx), float(fy), float(fz)


class BuoyancyTracer:
    """Buoyancy is the tracer `b`; `gravity_unit_vector` defaults to NegativeZDirection (0, 0, -1)."""

    def __init__(self, gravity_unit_vector=(0.0, 0.0, -1.0)):
        self.gravity_unit_vector = tuple(float(g) for g in gravity_unit_vector)


def get_coriolis_frequency_components(coriolis):
    """src/Oceanostics.jl:178-194."""
    if hasattr(coriolis, "coriolis") and not isinstance(coriolis, (FPlane, ConstantCartesianCoriolis)):
        coriolis = coriolis.coriolis       # a model
    if isinstance(coriolis, FPlane):
        return 0.0, 0.0, coriolis.f
    if isinstance(coriolis, ConstantCartesianCoriolis):
        return coriolis.fx, coriolis.fy, coriolis.fz
    if coriolis is None:
        return 0.0, 0.0, 0.0
    raise ValueError("Extraction of rotation components is only implemented for `FPlane`, "
                     "`ConstantCartesianCoriolis` and `nothing`.")


class Forcing:
    """Oceananigans `Forcing(func; field_dependencies, parameters)` in continuous form: `func(x, y, z, t, deps..., [p])`
    evaluated at the location of the forced field.  The host layer materialises it (`ForcingField`) the way Oceananigans
    evaluates the functor inside its kernels: at every stored index, halos included, from the dependency's own values there.
    `func` must be written with array operations (it receives torch tensors)."""

    def __init__(self, func, field_dependencies=(), parameters=None):
        self.func = func
        self.field_dependencies = (field_dependencies,) if isinstance(field_dependencies, str) else tuple(field_dependencies)
        self.parameters = parameters


class NonhydrostaticModel:
    hydrostatic = False

    def __init__(self, grid, tracers=("b",), closure=None, coriolis=None, buoyancy=None, advection="Centered",
                 closure_fields=None, forcing=None):
        self.grid = grid
        self.velocities = SimpleNamespace(u=XFaceField(grid), v=YFaceField(grid), w=ZFaceField(grid))
        self.tracers = SimpleNamespace(**{name: CenterField(grid) for name in tracers})
        self.pressures = SimpleNamespace(pNHS=CenterField(grid), pHY=None)
        self.closure = closure if closure is not None else Closure()
        self.closure_fields = closure_fields
        self.coriolis = coriolis
        self.buoyancy = buoyancy if buoyancy is not None else (BuoyancyTracer() if "b" in tracers else None)
        self.advection = advection
        self.clock = SimpleNamespace(time=0.0, iteration=0)
        self.velocities.w.boundary["impenetrable"] = True
        forcing = dict(forcing or {})
        unknown = set(forcing) - {"u", "v", "w"} - set(tracers)
        if unknown:
            raise ValueError(f"forcing names {sorted(unknown)} are not fields of the model")
        self.forcing = SimpleNamespace(**{name: forcing.get(name) for name in ("u", "v", "w") + tuple(tracers)})
        self.stokes_drift = self.background_fields = self.immersed_boundary = None
        if hasattr(self.closure, "bind"):          # closures whose eddy fields the library computes (SmagorinskyLilly)
            self.closure.bind(self)

    def compute_diffusivities(self, time=None):
        """Oceananigans' `compute_diffusivities!` for the closures the library evaluates: refresh model.closure_fields."""
        from .operations import compute
        for f in self.closure.field_arguments():
            if getattr(f, "is_computed", False):
                compute(f, time)
        return self.closure_fields

    def fields(self):
        return {**vars(self.velocities), **vars(self.tracers)}

    def fill_halo_regions(self):
        for f in list(vars(self.velocities).values()) + list(vars(self.tracers).values()) + [self.pressures.pNHS]:
            if hasattr(f, "fill_halo_regions"):
                f.fill_halo_regions()
        for f in self.closure.field_arguments():
            f.fill_halo_regions()
        return self


class HydrostaticFreeSurfaceModel(NonhydrostaticModel):
    """For the diagnostics the only difference is that perturbation_fields substitutes w = ZeroField()
    (src/Oceanostics.jl:160-162)."""
    hydrostatic = True


def validate_location(location, name, valid=("c", "c", "c")):
    """src/Oceanostics.jl:114-116."""
    from .fields import normalize_location
    if normalize_location(location) != tuple(valid):
        raise ValueError(f"{name} only supports location = {valid} for now.")


def perturbation_velocities(model, U=None, V=None, W=None):
    """perturbation_fields(model; u=U, v=V, w=W) (src/Oceanostics.jl:157-173) as operand slots: the resolved
    velocities plus the means to subtract (evaluated on the fly, never materialised)."""
    from .fields import as_operand
    u, v, w = model.velocities.u, model.velocities.v, model.velocities.w
    if model.hydrostatic:
        w = ZeroField()
    return dict(u=u, v=v, w=w, U=as_operand(U), V=as_operand(V), W=as_operand(W))


def advection_flags(model, what):
    """Request flags for the model's advection scheme: Centered second order (0) or Centered(order = 4) (OCB_REQ_CENTERED4) -- the
    schemes every example of the reference uses (docs/examples/*.jl: `advection = Centered(order=4)`); anything else (upwind-biased,
    WENO) is refused: its reconstruction is not restated."""
    from . import _lib
    adv = str(getattr(model, "advection", "Centered")).replace(" ", "")
    if adv in ("Centered", "Centered(order=2)", "Centered()"):
        return 0
    if adv == "Centered(order=4)":
        return _lib.OCB_REQ_CENTERED4
    raise _lib.OcbError(f"{what}: advection scheme {model.advection!r} -- only Centered(order = 2 | 4) advection is evaluated by the CUDA library")
