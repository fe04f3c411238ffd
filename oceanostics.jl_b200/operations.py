"""KernelFunctionOperation / computed Field / Integral / Average, and the multi-diagnostic fusion planner.

Mirrors the Oceananigans seam the reference plugs into:
  * constructors return `KernelFunctionOperation`s (reference: src/KineticEnergyEquation.jl:491-492 and every
    other constructor);
  * `Field(op)` makes a computed field, `compute(field)` (`compute!`) evaluates it: refresh Field arguments
    (`compute_at!`), write only the destination's index window, `fill_halo_regions!`, stamp the status
    (the recipe of src/SpatialFilters/SpatialFilters.jl:347-378);
  * `Integral(op)` / `Average(op)` / `∫dV` reduce without materialising the operand;
  * `FusedDiagnostics(...)` is the sibling registry: the first `compute` at a new `time` runs ONE sweep for
    every member and stamps them all (the delegation pattern of src/BackgroundPotentialEnergyEquation.jl:246-262).

Everything that touches field values runs in liboceanostics_b200 on the GPU; there is no CPU route.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from .fields import (AbstractField, BinaryOperation, ConstantField, Field, ZeroField, as_operand, current_stream_ptr,
                     normalize_location)
from .grids import Center, Face

_SLOTS = ("u", "v", "w", "b", "p", "U", "V", "W", "B")


class SweepOperands:
    """The operands one diagnostic reads: field slots, closure, scalar parameters."""

    def __init__(self, **kw):
        self.fields = {}     # slot name -> operand (Field / ZeroField / ConstantField)
        self.aux = {}        # aux index -> Field
        self.closure = None  # Closure or None
        self.scalars = {}    # "f", "dir", "f_dir", "ghat", "ro_bg", "les"
        for k, v in kw.items():
            if k in _SLOTS:
                self.fields[k] = as_operand(v)
            elif k == "aux":
                self.aux = dict(v)
            elif k == "closure":
                self.closure = v
            else:
                self.scalars[k] = v

    def merged_with(self, other: "SweepOperands"):
        """Union of two operand sets, or None when they conflict (then the two cannot share a sweep)."""
        out = SweepOperands()
        out.fields = dict(self.fields)
        for k, v in other.fields.items():
            if k in out.fields and not _same_operand(out.fields[k], v):
                return None
            out.fields[k] = v
        out.aux = dict(self.aux)
        for k, v in other.aux.items():
            if k in out.aux and out.aux[k] is not v:
                return None
            out.aux[k] = v
        if self.closure is not None and other.closure is not None and self.closure is not other.closure \
                and not self.closure.same_as(other.closure):
            return None
        out.closure = self.closure if self.closure is not None else other.closure
        out.scalars = dict(self.scalars)
        for k, v in other.scalars.items():
            if k in out.scalars and tuple(np.atleast_1d(out.scalars[k])) != tuple(np.atleast_1d(v)):
                return None
            out.scalars[k] = v
        return out

    def descriptor(self) -> _lib.ocb_sweep_inputs:
        s = _lib.ocb_sweep_inputs()
        zero = ZeroField().descriptor()
        for name in _SLOTS:
            setattr(s, name, self.fields[name].descriptor() if name in self.fields else zero)
        for a in range(_lib.OCB_N_AUX):
            s.aux[a] = self.aux[a].descriptor() if a in self.aux else zero
        if self.closure is not None:
            s.closure = self.closure.descriptor()
        for d in range(3):
            s.f[d] = float(self.scalars.get("f", (0.0, 0.0, 0.0))[d])
            s.dir[d] = float(self.scalars.get("dir", (0.0, 0.0, 1.0))[d])
            s.ghat[d] = float(self.scalars.get("ghat", (0.0, 0.0, 1.0))[d])
        s.f_dir = float(self.scalars.get("f_dir", 0.0))
        for d in range(6):
            s.ro_bg[d] = float(self.scalars.get("ro_bg", (0.0,) * 6)[d])
        for d in range(4):
            s.les[d] = float(self.scalars.get("les", (0.0,) * 4)[d])
        return s

    def field_arguments(self):
        out = [f for f in self.fields.values() if isinstance(f, Field)]
        out += [f for f in self.aux.values() if isinstance(f, Field)]
        if self.closure is not None:
            out += self.closure.field_arguments()
        return out


def _same_operand(a, b):
    if a is b:
        return True
    if isinstance(a, ZeroField) and isinstance(b, ZeroField):
        return True
    if isinstance(a, ConstantField) and isinstance(b, ConstantField):
        return a.value == b.value
    return False


class Closure:
    """Isotropic scalar-diffusivity closure(s) as the library sees them (`ocb_closure`): a constant ν/κ
    (ScalarDiffusivity) plus up to two eddy-viscosity / diffusivity ccc fields (SmagorinskyLilly,
    AnisotropicMinimumDissipation); a closure tuple is the sum.  Anisotropic formulations are rejected as
    the reference does (validate_dissipative_closure, src/Oceanostics.jl:118-120)."""

    def __init__(self, nu=0.0, kappa=0.0, nu_fields=(), kappa_fields=(), kappa_scale=None, isotropic=True, name="ScalarDiffusivity"):
        if not isotropic:
            raise ValueError(f"Closure {name} is not isotropic; only isotropic (ThreeDimensionalFormulation) closures are supported")
        self.nu, self.kappa = float(nu), float(kappa)
        self.nu_fields, self.kappa_fields = tuple(nu_fields), tuple(kappa_fields)
        self.kappa_scale = tuple(kappa_scale) if kappa_scale is not None else (1.0,) * len(self.kappa_fields)
        self.name = name
        if len(self.nu_fields) > _lib.OCB_MAX_CLOSURE_FIELDS or len(self.kappa_fields) > _lib.OCB_MAX_CLOSURE_FIELDS:
            raise ValueError(f"at most {_lib.OCB_MAX_CLOSURE_FIELDS} eddy fields per closure tuple")

    @staticmethod
    def tuple_of(*closures):
        """A closure tuple: fluxes (hence viscosities) add (src/Oceanostics.jl:201-208)."""
        nu = sum(c.nu for c in closures)
        ka = sum(c.kappa for c in closures)
        nf = tuple(f for c in closures for f in c.nu_fields)
        kf = tuple(f for c in closures for f in c.kappa_fields)
        ks = tuple(s for c in closures for s in c.kappa_scale)
        return Closure(nu, ka, nf, kf, ks, name="(" + ", ".join(c.name for c in closures) + ")")

    def same_as(self, o):
        return (self.nu == o.nu and self.kappa == o.kappa and len(self.nu_fields) == len(o.nu_fields)
                and all(a is b for a, b in zip(self.nu_fields, o.nu_fields))
                and len(self.kappa_fields) == len(o.kappa_fields)
                and all(a is b for a, b in zip(self.kappa_fields, o.kappa_fields)) and self.kappa_scale == o.kappa_scale)

    def field_arguments(self):
        return list(self.nu_fields) + list(self.kappa_fields)

    def descriptor(self):
        c = _lib.ocb_closure()
        c.nu_const, c.kappa_const = self.nu, self.kappa
        c.n_nu_fields, c.n_kappa_fields = len(self.nu_fields), len(self.kappa_fields)
        for i, f in enumerate(self.nu_fields):
            c.nu_fields[i] = f.descriptor()
        for i, f in enumerate(self.kappa_fields):
            c.kappa_fields[i] = f.descriptor()
            c.kappa_scale[i] = self.kappa_scale[i]
        return c


class SmagorinskyLilly(Closure):
    """Oceananigans' `SmagorinskyLilly(C, Cb, Pr)` with its eddy viscosity computed ON THE DEVICE (SURVEY.md section 8f row 3:
    the `compute_diffusivities!` step before the path).  `bind(model)` -- called by the model constructor -- makes
    `model.closure_fields.νₑ` a computed Field of the library's OCB_DIAG_NU_SMAG kernel over the model's velocities and
    buoyancy; every diagnostic built with this closure refreshes it first (`compute_at!`), so no νₑ array is handed in.
    κₑ = νₑ / Pr.  `Cb = None` (Oceananigans' default): no stratification correction."""

    def __init__(self, C=0.16, Cb=None, Pr=1.0):
        Closure.__init__(self, name="SmagorinskyLilly")
        self.C, self.Cb, self.Pr = float(C), (None if Cb is None else float(Cb)), float(Pr)
        self.nu_e = None

    def bind(self, model):
        from types import SimpleNamespace
        v = model.velocities
        kw = dict(u=v.u, v=v.v, w=v.w, les=(self.C, self.Cb or 0.0, self.Pr, 0.0))
        if self.Cb:
            kw["b"] = model.tracers.b
        op = KernelFunctionOperation("NU_SMAG", model.grid, SweepOperands(**kw), name="SmagorinskyLillyViscosity",
                                     computes="eddy viscosity  νₑ = ς (C Δᶠ)² √(2 ΣᵢⱼΣᵢⱼ)")
        self.nu_e = ComputedField(op)
        self.nu_fields, self.kappa_fields, self.kappa_scale = (self.nu_e,), (self.nu_e,), (1.0 / self.Pr,)
        model.closure_fields = SimpleNamespace(nu_e=self.nu_e, **{"νₑ": self.nu_e})
        return self


class AbstractOperation(AbstractField):
    pass


class KernelFunctionOperation(AbstractOperation):
    """`KernelFunctionOperation{LX,LY,LZ}(kernel_function, grid, args...)`: `diag` names the kernel function
    (an OCB_DIAG_* id), `operands` are its arguments."""

    def __init__(self, diag: str, grid, operands: SweepOperands, flags=0, name=None, computes=None):
        self.diag, self.grid, self.operands, self.flags = diag, grid, operands, int(flags)
        self.location = tuple(_lib.DIAG_LOC[diag])
        self.name = name or diag
        self.computes = computes

    def _as_computed_field(self, indices=None):
        return ComputedField(self, indices=indices)

    def __repr__(self):
        loc = ", ".join({"c": "Center", "f": "Face"}[l] for l in self.location)
        s = f"{self.name} KernelFunctionOperation at ({loc})\n├── grid: {self.grid!r}\n└── kernel id: OCB_DIAG_{self.diag}"
        if self.computes:
            s += f"\n└── computes: {self.computes}"
        return s


class ComputedField(Field):
    """`Field(op)` for a KernelFunctionOperation."""

    def __new__(cls, *a, **k):
        return object.__new__(cls)

    is_computed = True

    def __init__(self, operand: KernelFunctionOperation, indices=None):
        if getattr(self, "_initialized", False):
            return   # Field(op) re-enters __init__ after __new__ handed back the finished object
        self._initialized = True
        Field.__init__(self, operand.location, operand.grid, indices=indices)
        self.operand = operand
        self.group = None
        self.boundary["impenetrable"] = False   # computed fields carry default (no-flux / periodic) halos

    def compute(self, time=None):
        return compute(self, time)


class Reduction(AbstractOperation):
    """`Integral(op)` / `Average(op)` over the whole volume (`dims=:`)."""

    def __init__(self, kind, operand, indices=None):
        if not isinstance(operand, KernelFunctionOperation):
            raise TypeError("Integral / Average take a KernelFunctionOperation of this library")
        self.kind, self.operand, self.grid = kind, operand, operand.grid
        # optional 1-based inclusive index window per axis (the reduction of a windowed view of the operand)
        self.window = None
        if indices is not None:
            ext = self.grid.extent_of(operand.location)
            self.window = [(1, ext[d]) if w is None else (int(w[0]), int(w[1])) for d, w in enumerate(indices)]

    def _as_computed_field(self):
        return ReducedField(self)


def Integral(op, indices=None):
    return Reduction("integral", op, indices)


def Average(op, indices=None):
    if indices is not None:
        raise ValueError("Average over an index window is not supported; use Integral(op, indices=...)")
    return Reduction("average", op)


class ReducedField:
    """`Field(Integral(op))`: a scalar living in a device double; `value()` brings it to the host."""

    def __init__(self, reduction: Reduction):
        self.reduction, self.operand, self.grid = reduction, reduction.operand, reduction.grid
        self.result = torch.zeros(1, dtype=torch.float64, device=self.grid.device)
        self.group = None
        self.status_time = None

    def compute(self, time=None):
        return compute(self, time)

    def value(self) -> float:
        v = float(self.result.item())
        if self.reduction.kind == "average":
            v /= self.grid.total_volume(self.operand.location)
        return v


# ---- plans ------------------------------------------------------------------------------------------------
class SweepPlan:
    """One `ocb_plan`: a set of (operation, destination field or None, integral slot or -1) evaluated in one sweep."""

    def __init__(self, grid, items):
        if grid.device.type != "cuda":
            raise _lib.OcbError("diagnostics are evaluated by the CUDA library only: the grid must live on a CUDA device "
                                "(no CPU fallback)")
        if not grid.sweep_supported():
            raise ValueError("the stencil sweep needs regularly spaced x and y axes (z may be stretched)")
        self.grid, self.items = grid, items
        self.operands, self._inputs, self._reqs, self.n_slots = assemble_sweep(items)
        self.integrals = torch.zeros(max(self.n_slots, 1), dtype=torch.float64, device=grid.device)
        reqs = self._reqs
        handle = C.c_void_p()
        gd = grid.descriptor()
        _lib.check(_lib.lib().ocb_plan_create(C.byref(gd), C.byref(self._inputs), reqs, len(items), C.byref(handle)))
        self.handle = handle
        nl, var, ab = C.c_int32(), C.c_int32(), C.c_double()
        _lib.check(_lib.lib().ocb_plan_info(handle, C.byref(nl), C.byref(var), C.byref(ab)))
        self.n_launches, self.variant, self.algorithmic_bytes = nl.value, var.value, ab.value

    def execute(self):
        _lib.check(_lib.lib().ocb_plan_execute(self.handle, self.integrals.data_ptr(), current_stream_ptr(self.grid.device)))

    @property
    def ring_capable(self):
        """True when `execute_ring` accepts this plan (integral-only class-sum sweep over a whole y-periodic slab)."""
        return bool(_lib.lib().ocb_plan_ring_capable(self.handle))

    def execute_ring(self, ring, step):
        """One slab's step of a peer-memory ring (`ocb_plan_execute_ring`): the sweep with its edge tile rows waiting for
        the neighbours' pushes of `step`, then the integrals summed over the ranks -- `self.integrals` holds the GLOBAL sums."""
        with torch.cuda.device(self.grid.device):
            _lib.check(_lib.lib().ocb_plan_execute_ring(self.handle, self.integrals.data_ptr(), C.byref(ring.desc), int(step),
                                                        current_stream_ptr(self.grid.device)))

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                _lib.lib().ocb_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


def assemble_sweep(items):
    """Merge the operands of `items` = [(operation, destination Field or None, integral slot or -1)] and build the
    C descriptors of one fused sweep.  Returns (operands, ocb_sweep_inputs, ocb_request array, number of slots)."""
    operands = SweepOperands()
    for item in items:
        operands = operands.merged_with(item[0].operands)
        if operands is None:
            raise ValueError("these diagnostics read conflicting operands and cannot share one sweep")
    n_slots = 1 + max([item[2] for item in items] + [-1])
    reqs = (_lib.ocb_request * len(items))()
    for r, item in enumerate(items):
        op, dest, slot = item[:3]
        fill_request(reqs[r], op, dest, slot, item[3] if len(item) > 3 else None)
    return operands, operands.descriptor(), reqs, n_slots


def fill_request(req, op: KernelFunctionOperation, dest, slot, window=None):
    if window is not None:
        for d in range(3):
            req.window_lo[d], req.window_hi[d] = window[d]
    req.diag = _lib.DIAG[op.diag]
    req.flags = op.flags
    req.integral_slot = slot
    req.out = None
    if dest is not None:
        req.out = dest.data.data_ptr()
        strides = (dest.data.stride(2), dest.data.stride(1), dest.data.stride(0))
        for d in range(3):
            req.out_halo[d] = dest.halo[d]
            req.out_stride[d] = strides[d]
            req.out_offset[d] = dest.window[d][0]
            if dest.windowed:
                req.window_lo[d], req.window_hi[d] = dest.window[d]


def _refresh_arguments(operands: SweepOperands, time):
    """`compute_at!` on every Field argument that is itself computed (status-gated)."""
    for f in operands.field_arguments():
        if getattr(f, "is_computed", False):
            compute_at(f, time)


def compute(field, time=None):
    """`compute!(field, time)`."""
    if getattr(field, "group", None) is not None:
        field.group.compute(time)
        return field
    if isinstance(field, ReducedField):
        if getattr(field, "_plan", None) is None:
            field._plan = SweepPlan(field.grid, [(field.operand, None, 0, field.reduction.window)])
        _refresh_arguments(field._plan.operands, time)
        field._plan.execute()
        field.result.copy_(field._plan.integrals[:1])
        field.status_time = time
        return field
    if isinstance(field, ComputedField):
        if getattr(field, "_plan", None) is None:
            field._plan = SweepPlan(field.grid, [(field.operand, field, -1)])
        _refresh_arguments(field._plan.operands, time)
        field._plan.execute()
        field.fill_halo_regions()
        field.status_time = time
        return field
    if hasattr(field, "compute"):
        return field.compute(time)
    return field


def _up_to_date(status_time, time):
    """Oceananigans' `compute_at!` rule: recompute when `time == zero(time) || time != status.time` -- so always at t = 0
    (fields set after a first evaluation at t = 0 must not leave stale diagnostics) and whenever no time is given."""
    return time is not None and time != 0 and status_time is not None and status_time == time


def compute_at(field, time):
    """`compute_at!`: recompute unless the status time says the field is current (never skipped at t = 0)."""
    if _up_to_date(getattr(field, "status_time", None), time):
        return field
    return compute(field, time)


class FusedDiagnostics:
    """Sibling registry: one sweep over u, v, w, b, p emits every member (computed fields and reductions)."""

    def __init__(self, *members):
        flat = []
        for m in members:
            flat.extend(m if isinstance(m, (list, tuple)) else [m])
        self.members = []
        for m in flat:
            if isinstance(m, (KernelFunctionOperation, Reduction)):
                m = m._as_computed_field()
            if not isinstance(m, (ComputedField, ReducedField)):
                raise TypeError("FusedDiagnostics takes operations, computed Fields or Integral/Average fields")
            self.members.append(m)
        if not self.members:
            raise ValueError("FusedDiagnostics needs at least one member")
        if len(self.members) > _lib.OCB_MAX_REQ:
            raise ValueError(f"at most {_lib.OCB_MAX_REQ} diagnostics per fused sweep")
        self.grid = self.members[0].grid
        items, slot = [], 0
        for m in self.members:
            if isinstance(m, ReducedField):
                items.append((m.operand, None, slot, m.reduction.window))
                m._slot = slot
                slot += 1
            else:
                items.append((m.operand, m, -1))
            m.group = self
        self.plan = SweepPlan(self.grid, items)
        self.status_time = None
        self._computed_once = False

    def compute(self, time=None):
        if self._computed_once and _up_to_date(self.status_time, time):
            return self
        _refresh_arguments(self.plan.operands, time)
        self.plan.execute()
        for m in self.members:
            if isinstance(m, ReducedField):
                m.result.copy_(self.plan.integrals[m._slot:m._slot + 1])
            else:
                m.fill_halo_regions()
            m.status_time = time
        self.status_time, self._computed_once = time, True
        return self

    def integrals(self):
        """Host copy of every reduction member's value (one D2H transfer)."""
        vals = self.plan.integrals.cpu().numpy()
        out = []
        for m in self.members:
            if isinstance(m, ReducedField):
                v = float(vals[m._slot])
                if m.reduction.kind == "average":
                    v /= self.grid.total_volume(m.operand.location)
                out.append(v)
        return out


def materialize_binary(op: BinaryOperation, location=None) -> Field:
    """`Field(a ∘ b)` through the library's pointwise kernel (operands interpolated to the result location)."""
    grid = op.grid
    if grid is None or grid.device.type != "cuda":
        raise _lib.OcbError("Field(a ∘ b) is evaluated by the CUDA library only (no CPU fallback)")
    a, b = op.a, op.b
    for x in (a, b):
        if isinstance(x, BinaryOperation):
            raise TypeError("nested lazy operations: materialise the inner one with Field(...) first")
    loc = normalize_location(location) if location is not None else op.location
    out = Field(loc, grid)
    out.boundary["impenetrable"] = False
    code = {"+": _lib.OCB_OP_ADD, "-": _lib.OCB_OP_SUB, "*": _lib.OCB_OP_MUL}[op.op]
    gd, da, db, dd = grid.descriptor(), a.descriptor(), b.descriptor(), out.descriptor()
    _lib.check(_lib.lib().ocb_pointwise(C.byref(gd), code, 1.0, C.byref(da), C.byref(db), C.byref(dd),
                                        current_stream_ptr(grid.device)))
    out.fill_halo_regions()
    return out


class PointwiseField(Field):
    """A computed `Field` for `@at loc a`, `a + b`, `a - b` or `a * b` of (possibly computed) Fields: recomputed by
    `compute` / `compute_at` like any other computed field (operands first), through the library's pointwise kernel,
    which interpolates each operand to the result's location with Oceananigans' two-point means."""
    is_computed = True

    def __new__(cls, *a, **k):
        return object.__new__(cls)

    def __init__(self, op, a, b=None, location=None):
        for x in (a, b):
            if isinstance(x, AbstractOperation) and not isinstance(x, Field):
                raise TypeError("materialise operations with Field(op) before combining them pointwise")
        grid = a.grid
        loc = normalize_location(location) if location is not None else a.location
        Field.__init__(self, loc, grid)
        self.op, self.a, self.b = op, a, (as_operand(b) if op != "@" else None)
        self.group = None
        self.boundary["impenetrable"] = False

    def compute(self, time=None):
        grid = self.grid
        if grid.device.type != "cuda":
            raise _lib.OcbError("pointwise fields are evaluated by the CUDA library only (no CPU fallback)")
        for x in (self.a, self.b):
            if getattr(x, "is_computed", False):
                compute_at(x, time)
        code = {"@": _lib.OCB_OP_COPY, "+": _lib.OCB_OP_ADD, "-": _lib.OCB_OP_SUB, "*": _lib.OCB_OP_MUL}[self.op]
        gd, da, dd = grid.descriptor(), self.a.descriptor(), self.descriptor()
        db = self.b.descriptor() if self.b is not None else None
        _lib.check(_lib.lib().ocb_pointwise(C.byref(gd), code, 1.0, C.byref(da), C.byref(db) if db is not None else None, C.byref(dd),
                                            current_stream_ptr(grid.device)))
        self.fill_halo_regions()
        self.status_time = time
        return self


class ForcingField(Field):
    """A model forcing evaluated at the points of the field it forces: a computed `Field` (recomputed by `compute` /
    `compute_at` like any other).  `forcing` is a `models.Forcing` (continuous form `func(x, y, z, t, deps..., [p])`), a
    number or a Field at the same location.  Dependencies must live at the forced field's location (the reference's own
    test forces each velocity with a function of itself, test/test_kinetic_energy_equation.jl:105-129); the functor runs
    as torch array operations on the device -- the step Oceananigans performs inside its tendency kernels."""
    is_computed = True

    def __new__(cls, *a, **k):
        return object.__new__(cls)

    def __init__(self, forcing, forced: Field, model):
        Field.__init__(self, forced.location, forced.grid)
        self.forcing, self.forced, self.model = forcing, forced, model
        self.group = None
        self.boundary["impenetrable"] = False

    def compute(self, time=None):
        import torch
        g, F = self.grid, self.forcing
        t = float(getattr(self.model.clock, "time", 0.0)) if time is None else float(time)
        shape = self.data.shape                              # (z, y, x) parent, halos included
        idx = [np.arange(1 - self.halo[d], self.stored[d] + self.halo[d] + 1) for d in range(3)]
        node = lambda d: torch.as_tensor(np.asarray(g.nodes(d, self.location[d], idx[d]), dtype=np.float64), device=self.data.device,
                                         dtype=self.data.dtype)
        x, y, z = node(0)[None, None, :], node(1)[None, :, None], node(2)[:, None, None]
        deps = []
        for name in F.field_dependencies:
            f = self.model.fields()[name]
            if f.location != self.location or tuple(f.data.shape) != tuple(shape):
                raise ValueError(f"forcing dependency `{name}` is not located where the forced field is")
            if getattr(f, "is_computed", False):
                compute_at(f, time)
            deps.append(f.data)
        args = [x, y, z, t] + deps + ([F.parameters] if F.parameters is not None else [])
        out = F.func(*args)
        self.data.copy_(torch.broadcast_to(torch.as_tensor(out, device=self.data.device, dtype=self.data.dtype), shape))
        self.status_time = time
        return self


def forcing_operand(forcing, forced, model):
    """`model.forcing.<name>` as an operand: nothing -> ZeroField, a number -> constant, a Field as is, a functor -> ForcingField."""
    from .models import Forcing
    if forcing is None:
        return ZeroField()
    if isinstance(forcing, Forcing):
        return ForcingField(forcing, forced, model)
    if isinstance(forcing, Field):
        if forcing.location != forced.location:
            raise ValueError("a forcing Field must be located where the field it forces is")
        return forcing
    return as_operand(forcing)


def perturbation_split(x, default_mean=None):
    """Pattern-match a velocity/tracer argument: `field`, or the lazy `field - mean` the reference's
    constructors build (src/TurbulentKineticEnergyEquation.jl:327-330).  Returns (field, mean or None)."""
    if isinstance(x, BinaryOperation) and x.op == "-":
        return x.a, x.b
    return x, default_mean
