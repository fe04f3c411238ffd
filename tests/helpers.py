"""Shared test plumbing: build the SAME synthetic state in the NumPy oracle and in the package, run a set of
diagnostics through (a) the CUDA library via its C ABI, or (b) the host-compiled evaluators
(tests/hostcheck, CPU container only), and compare with the oracle.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import the oracle.
"""
from __future__ import annotations

import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import ocb_loader  # noqa: E402

ocb = ocb_loader.load()
from oracle.grid import OGrid, OField, ConstantField as OConst, ZeroField as OZero, DiffField, average_dims  # noqa: E402
from oracle import kernels as K  # noqa: E402

SEED = 20261019
_TOPO = {"P": ocb.Periodic, "B": ocb.Bounded, "F": ocb.Flat}


# ---- grids -----------------------------------------------------------------------------------------------
def stretched_faces(N):
    """The reference's stretched test grid (test/test_utils.jl:17-24)."""
    S = 0.99
    f_asin = lambda k: -np.arcsin(S * (2 * k - N - 2) / N) / np.pi + 0.5
    F1, F2 = f_asin(1), f_asin(N + 1)
    return lambda k: ((F1 + F2) / 2 - f_asin(k)) / (F1 - F2)


def make_grids(size, stretched=False, dtype=np.float64, device="cpu", halo=(3, 3, 3), topology=("P", "P", "B"), extent=(1, 1, 1)):
    """-> (oracle grid, package grid) describing the same domain."""
    tdtype = torch.float64 if dtype == np.float64 else torch.float32
    topo = tuple(_TOPO[t] for t in topology)
    if stretched:
        zf = stretched_faces(size[2])
        og = OGrid(size, x=(0, extent[0]), y=(0, extent[1]), z=zf, halo=halo, topology=topology, dtype=dtype)
        pg = ocb.RectilinearGrid(size, x=(0, extent[0]), y=(0, extent[1]), z=zf, halo=halo, topology=topo, dtype=tdtype, device=device)
    else:
        og = OGrid(size, extent=extent, halo=halo, topology=topology, dtype=dtype)
        pg = ocb.RectilinearGrid(size, extent=extent, halo=halo, topology=topo, dtype=tdtype, device=device)
    return og, pg


def mirror(ofield: OField, pgrid, impenetrable=False):
    """Package Field holding exactly the oracle field's parent array (halos included)."""
    loc = tuple(ofield.loc)
    f = ocb.Field(loc, pgrid)
    arr = np.ascontiguousarray(ofield.data.transpose(2, 1, 0))
    assert tuple(arr.shape) == tuple(f.data.shape), (arr.shape, f.data.shape)
    f.data.copy_(torch.from_numpy(arr))
    f.boundary["impenetrable"] = impenetrable
    return f


# ---- synthetic state ---------------------------------------------------------------
class State:
    """u, v, w, b, p (+ eddy viscosity) on both sides, halos filled by the oracle's fill_halo."""

    def __init__(self, size=(16, 12, 10), stretched=False, dtype=np.float64, device="cpu", seed=SEED, topology=("P", "P", "B"),
                 extent=(1, 1, 1), amplitude=1.0, halo=(3, 3, 3)):
        self.og, self.pg = make_grids(size, stretched, dtype, device, halo=halo, topology=topology, extent=extent)
        rng = np.random.default_rng(seed)
        og = self.og
        two_pi = 2 * np.pi

        def fld(loc, fn, noise, impenetrable=True):
            f = OField(og, loc)
            f.set(fn)
            f.interior[...] += (noise * rng.standard_normal(f.interior.shape)).astype(og.dtype)
            return f.fill_halo(impenetrable=impenetrable)

        a = amplitude
        N2 = 1e-5
        self.ou = fld("fcc", lambda x, y, z: a * np.sin(two_pi * x) * np.cos(two_pi * y), 0.1 * a)
        self.ov = fld("cfc", lambda x, y, z: -a * np.cos(two_pi * x) * np.sin(two_pi * y), 0.1 * a)
        self.ow = fld("ccf", lambda x, y, z: 0.0, 0.1 * a)
        self.ob = fld("ccc", lambda x, y, z: N2 * z, 0.1 * N2 * extent[2])
        self.op = fld("ccc", lambda x, y, z: 0.0, 0.1)
        nu = OField(og, "ccc")
        nu.interior[...] = (1e-3 * (1.0 + 0.5 * np.sin(two_pi * og.node(0, "c", og.indices("ccc")[0])) +
                                    0.25 * rng.random(nu.interior.shape))).astype(og.dtype)
        self.onu = nu.fill_halo()
        pg = self.pg
        self.u, self.v = mirror(self.ou, pg), mirror(self.ov, pg)
        self.w = mirror(self.ow, pg, impenetrable=True)
        self.b, self.p, self.nu = mirror(self.ob, pg), mirror(self.op, pg), mirror(self.onu, pg)

    # closures on both sides: name -> (oracle closure, package closure)
    def closures(self):
        return {
            "scalar": (K.Closure(nu=1e-6, kappa=1e-7), ocb.Closure(nu=1e-6, kappa=1e-7)),
            "smagorinsky_like": (K.Closure(nu=self.onu, Pr=1.0),
                                 ocb.Closure(nu_fields=(self.nu,), kappa_fields=(self.nu,), kappa_scale=(1.0,), name="SmagorinskyLilly")),
            "tuple": ((K.Closure(nu=1e-6, kappa=1e-7), K.Closure(nu=self.onu, kappa=self.onu)),
                      ocb.Closure.tuple_of(ocb.Closure(nu=1e-6, kappa=1e-7),
                                           ocb.Closure(nu_fields=(self.nu,), kappa_fields=(self.nu,), name="AMD"))),
        }

    def model(self, closure=None, coriolis=None, buoyancy=None):
        m = ocb.NonhydrostaticModel(self.pg, tracers=("b",), closure=closure, coriolis=coriolis, buoyancy=buoyancy)
        m.velocities.u, m.velocities.v, m.velocities.w = self.u, self.v, self.w
        m.tracers.b = self.b
        m.pressures.pNHS = self.p
        if hasattr(m.closure, "bind"):
            m.closure.bind(m)            # eddy fields computed by the library read the fields just swapped in
        return m

    def ofields(self):
        return {"u": self.ou, "v": self.ov, "w": self.ow, "b": self.ob}


# ---- running a set of operations ---------------------------------------------------------------------------
_hc = None


def hostcheck_lib():
    global _hc
    if _hc is None:
        path = os.path.join(ROOT, "tests", "hostcheck", "libhostcheck.so")
        if not os.path.exists(path):
            import __graft_entry__ as ge
            ge.build_hostcheck()
        _hc = C.CDLL(path)
        _hc.hc_last_error.restype = C.c_char_p
        _hc.hc_sweep.restype = C.c_int
    return _hc


def run_ops(ops, grid, integrals=False, backend=None, fused=True):
    """Evaluate KernelFunctionOperations `ops` -> (list of interior arrays [i,j,k], list of integrals or None).

    backend 'cuda': through ocb_plan_* (one fused sweep when `fused`); backend 'host': the same descriptors handed to
    tests/hostcheck (CPU container)."""
    backend = backend or ("cuda" if grid.device.type == "cuda" else "host")
    if backend == "cuda":
        fields = [ocb.Field(op) for op in ops]
        reds = [ocb.Field(ocb.Integral(op)) for op in ops] if integrals else []
        if fused:
            # greedy grouping: a member joins the current sweep while its operands are compatible and there is room
            groups, cur, cur_ops = [], [], None
            for mbr in fields + reds:
                merged = mbr.operand.operands if cur_ops is None else cur_ops.merged_with(mbr.operand.operands)
                if merged is None or len(cur) == 16:
                    groups.append(cur)
                    cur, merged = [], mbr.operand.operands
                cur.append(mbr)
                cur_ops = merged
            groups.append(cur)
            for g in groups:
                ocb.FusedDiagnostics(*g).compute()
        else:
            for f in fields + reds:
                f.compute()
        torch.cuda.synchronize()
        return [f.interior_ijk() for f in fields], ([r.value() for r in reds] if integrals else None)
    from oceanostics_b200.operations import assemble_sweep
    outs = [ocb.Field(op.location, grid) for op in ops]
    vals, ints = [], []
    for n, (op, dest) in enumerate(zip(ops, outs)):
        items = [(op, dest, 0 if integrals else -1)]
        operands, inputs, reqs, n_slots = assemble_sweep(items)
        res = (C.c_double * 1)()
        gd = grid.descriptor()
        rc = hostcheck_lib().hc_sweep(C.byref(gd), C.byref(inputs), reqs, 1, res)
        if rc != 0:
            raise ValueError(hostcheck_lib().hc_last_error().decode())
        vals.append(dest.interior_ijk())
        ints.append(res[0])
    return vals, (ints if integrals else None)


def assert_close(got, want, rtol, what="", scale=None):
    """|got - want| <= rtol * max(|want|) elementwise: signed transport terms cancel to ~0 locally, so the
    tolerance is relative to the field's magnitude (SURVEY.md section 7 'reduction determinism & cancellation')."""
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, (what, got.shape, want.shape)
    s = scale if scale is not None else max(float(np.max(np.abs(want))), np.finfo(np.float64).tiny)
    err = float(np.max(np.abs(got - want)))
    assert err <= rtol * s, f"{what}: max abs err {err:.3e} > {rtol:g} * {s:.3e}"


def smoke_check(device="cuda:0"):
    """__graft_entry__.smoke(): one small invocation of each kernel family of the hot path on `device`, each checked against
    the oracle -- (1) field-writing budget kernel + generic kernel (eps, Ri fields and integrals), (2) the class-sum kernel
    (integrals-only budget plan) and its ring step, (3) the velocity-gradient kernel (eddy-viscosity closure, eps + |S|),
    (4) a 3-pass Gaussian filter, (5) the sort-based reference height."""
    from cases import check_case, integral_scale      # conditioning-aware for the Richardson number
    st = State(size=(32, 24, 16), device=device)
    oc, pc = st.closures()["scalar"]
    m = st.model(closure=pc)
    KE, FD, TV = ocb.KineticEnergyEquation, ocb.FlowDiagnostics, ocb.TracerVarianceEquation
    # ---- (1) fields + integrals: ocb_budget_kernel (OUT) and the generic kernel -------------------------------------
    ops = [KE.DissipationRate(m), FD.RichardsonNumber(m), KE.KineticEnergy(m), KE.KineticEnergyAdvection(m),
           KE.KineticEnergyPressureRedistribution(m), TV.TracerVarianceDissipationRate(m, "b")]
    vals, ints = run_ops(ops, st.pg, integrals=True, backend="cuda")
    og, f = st.og, st.ofields()
    par = {"closure": oc}
    want = [K.evaluate(K.viscous_dissipation_rate_ccc, og, "ccc", None, f, par),
            K.evaluate(K.richardson_number_ccf, og, "ccf", st.ou, st.ov, st.ow, st.ob, (0, 0, 1)),
            K.evaluate(K.kinetic_energy_ccc, og, "ccc", st.ou, st.ov, st.ow),
            K.evaluate(K.ke_advection_ccc, og, "ccc", f),
            K.evaluate(K.ke_pressure_ccc, og, "ccc", f, st.op),
            K.evaluate(K.tracer_variance_dissipation_rate_ccc, og, "ccc", oc, None, None, st.ob)]
    locs = ["ccc", "ccf", "ccc", "ccc", "ccc", "ccc"]
    for g, w_, loc, op in zip(vals, want, locs, ops):
        check_case(st, op.name, g, w_, 1e-12)
    for g, w_, loc, op in zip(ints, want, locs, ops):
        if op.name == "RichardsonNumber":
            continue
        assert abs(g - K.integral(w_, og, loc)) <= 1e-12 * integral_scale(og, w_, loc), (op.name, g)
    # ---- (2) integrals only: ocb_budget_sums_kernel, plain and as a one-rank ring step --------------------------------
    bops = [KE.KineticEnergy(m), KE.KineticEnergyAdvection(m), KE.KineticEnergyPressureRedistribution(m), KE.PotentialEnergyConversion(m),
            KE.DissipationRate(m), TV.TracerVarianceDissipationRate(m, "b"), TV.TracerVarianceDiffusion(m, "b")]
    bwant = [want[2], want[3], want[4], K.evaluate(K.ke_buoyancy_ccc, og, "ccc", f, (0, 0, 1), st.ob), want[0], want[5],
             K.evaluate(K.c_div_q_c, og, "ccc", oc, None, None, st.ob)]
    grp = ocb.FusedDiagnostics(*[ocb.Integral(o) for o in bops]).compute()
    assert grp.plan.variant == 1 and grp.plan.ring_capable
    from oceanostics_b200.distributed import RingSlabBudget
    ring = RingSlabBudget(st.pg, [st.u, st.v, st.w, st.b, st.p], lambda: bops, rank=0, world=1, timeout=5.0)
    ring_ints = ring.step().tolist()
    ring.check()
    for o, w_, a, b in zip(bops, bwant, grp.integrals(), ring_ints):
        ref, sc = K.integral(w_, og, "ccc"), integral_scale(og, w_, "ccc")
        assert abs(a - ref) <= 1e-12 * sc and abs(b - ref) <= 1e-12 * sc, (o.name, a, b, ref)
    # ---- (3) velocity-gradient kernel: eddy-viscosity closure, eps + |S| fields ---------------------------------------
    from cases import build_cases
    cases = {c[0]: c for c in build_cases(st, "smagorinsky_like")}
    names = ["DissipationRate", "StrainRateTensorModulus"]
    g3 = ocb.FusedDiagnostics(*[ocb.Field(cases[n][1]) for n in names]).compute()
    assert g3.plan.variant == 3, g3.plan.variant
    for n, mem in zip(names, g3.members):
        check_case(st, n, mem.interior_ijk(), cases[n][2])
    # ---- (4) staged Gaussian filter, three passes, shrink policy in z ------------------------------------------------
    from oracle import filters as OF
    got = ocb.Field(ocb.SpatialFilters.GaussianFilter(st.b, dims=(1, 2, 3), sigma=2.0 / 32, N=5)).compute().interior_ijk()
    np.testing.assert_allclose(got, OF.gaussian_filter(st.ob, (1, 2, 3), 2.0 / 32, 5), rtol=1e-13, atol=1e-18)
    # ---- (5) sort-based reference height (radix sort + scans + scatter) -----------------------------------------------
    from oracle import bpe as OB
    z = ocb.BackgroundPotentialEnergyEquation.reference_height(st.b)
    wantz = OB.reference_height(st.ob)
    torch.cuda.synchronize()
    n = got.size
    np.testing.assert_array_equal(z.sorted_buoyancy.cpu().numpy(), wantz["b_sorted"])
    np.testing.assert_allclose(z.interior_ijk(), wantz["zstar"], rtol=0, atol=max(1e-12, 4 * n * np.finfo(np.float64).eps))
