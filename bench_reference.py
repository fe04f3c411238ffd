"""CPU arm of bench.py: the oracle's C/OpenMP port (oracle/c/oracle_budget.c) timed on the host cores.

The reference itself is Julia + Oceananigans and cannot run in this image (no Julia), so the
"reference" arm is the declared stand-in the task contract allows: the per-cell, unshared, one-sweep-per-diagnostic
restatement of the reference's kernel functions, on all host threads, on a bounded sample of the same workload.
"""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
METRIC = "Gcells/s, fused KE + tracer-variance budget sweep (Float64)"


def _host_state(n, seed=20261019):
    """SURVEY.md section 8d synthetic fields on an n^3 sample (Periodic, Periodic, Bounded; halo 3), halos filled."""
    from oracle import cbudget
    H = 3
    rng = np.random.default_rng(seed)
    two_pi = 2 * np.pi
    xf = np.arange(n) / n
    xc = xf + 0.5 / n
    zc = -1 + (np.arange(n) + 0.5) / n

    def parent(nz):
        return np.zeros((nz + 2 * H, n + 2 * H, n + 2 * H))

    def fill(a, nz, center_z=True, wall_normal=False):
        it = a[H:H + nz, H:H + n, H:H + n]
        for m in range(H):   # periodic x then y
            a[:, :, H - 1 - m] = a[:, :, H + n - 1 - m]; a[:, :, H + n + m] = a[:, :, H + m]
        for m in range(H):
            a[:, H - 1 - m, :] = a[:, H + n - 1 - m, :]; a[:, H + n + m, :] = a[:, H + m, :]
        if wall_normal:
            a[H] = 0.0; a[H + nz - 1] = 0.0
        else:
            a[H - 1] = a[H]; a[H + nz] = a[H + nz - 1]
        return a

    u, v, w, b, p = parent(n), parent(n), parent(n + 1), parent(n), parent(n)
    u[H:H + n, H:H + n, H:H + n] = np.sin(two_pi * xf)[None, None, :] * np.cos(two_pi * xc)[None, :, None] + 0.1 * rng.standard_normal((n, n, n))
    v[H:H + n, H:H + n, H:H + n] = -np.cos(two_pi * xc)[None, None, :] * np.sin(two_pi * xf)[None, :, None] + 0.1 * rng.standard_normal((n, n, n))
    w[H:H + n + 1, H:H + n, H:H + n] = 0.1 * rng.standard_normal((n + 1, n, n))
    b[H:H + n, H:H + n, H:H + n] = 1e-5 * zc[:, None, None] + 1e-6 * rng.standard_normal((n, n, n))
    p[H:H + n, H:H + n, H:H + n] = 0.1 * rng.standard_normal((n, n, n))
    fill(u, n); fill(v, n); fill(w, n + 1, wall_normal=True); fill(b, n); fill(p, n)
    d = (1.0 / n, 1.0 / n, 1.0 / n)
    return cbudget.BudgetState((n, n, n), H, d, 1e-6, 1e-7, u, v, w, b, p)


def _time_budget(state, repeats):
    from oracle import cbudget
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        vals = [state.integral(d) for d in cbudget.DIAGS]     # one sweep per integral, as the reference does
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return best, vals


def cpu_baseline(args):
    """bench.py's `cpu_baseline` object: the port timed on a bounded sample (rank 0, N=1 only)."""
    from oracle import cbudget
    import __graft_entry__ as ge
    ge.build_oracle()
    n = int(args.cpu_size)
    st = _host_state(n)
    _time_budget(st, 1)            # warm-up (page faults, OpenMP team start)
    best, _ = _time_budget(st, 3)
    cores = cbudget.lib().ob_num_threads()
    return {"value": round(n ** 3 / best / 1e9, 5), "unit": "Gcells/s", "cores": cores, "kind": "port",
            "sample": f"{n}^3 Float64 sample of the same workload; 7 budget integrals, one OpenMP sweep per integral "
                      f"(the reference does not fuse), best of 3 after one warm-up; {best:.3f} s per step"}


def run_reference(args):
    """`bench.py --impl reference`: same metric/config line, timed on the host cores (rank 0 only)."""
    if int(os.environ.get("RANK", 0)) != 0:
        return
    from oracle import cbudget
    import __graft_entry__ as ge
    ge.build_oracle()
    n = int(args.cpu_size)
    st = _host_state(n)
    steps, warmup = max(1, min(args.steps, 5)), max(1, min(args.warmup, 1))
    for _ in range(warmup):
        _time_budget(st, 1)
    t0 = time.perf_counter()
    for _ in range(steps):
        _, vals = _time_budget(st, 1)
    dt = (time.perf_counter() - t0) / steps
    cores = cbudget.lib().ob_num_threads()
    value = n ** 3 / dt / 1e9
    size = tuple(args.size) if args.size else (1024, 1024, 1024)
    out = {"impl": "reference", "metric": METRIC, "value": round(value, 5), "unit": "Gcells/s", "n_gpus": args.gpus, "steps": steps,
           "warmup": warmup, "ms_per_step": round(dt * 1e3, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f64", "data": "synthetic",
           "config": {"workload": f"config 3A: {size[0]}x{size[1]}x{size[2]} Float64, fused KE budget {{K, ADV(Centered-2), PR, u_i b_i, eps}} "
                                  "+ tracer variance {chi, 2c div F}, 7 volume integrals",
                      "sample": f"each step = the 7 integrals on a {n}^3 sample of that workload (bounded so the run ends in minutes)"},
           "cpu_baseline": {"value": round(value, 5), "unit": "Gcells/s", "cores": cores, "kind": "port",
                            "sample": f"{n}^3 sample, one OpenMP sweep per integral, {cores} threads; Julia/Oceananigans are not "
                                      "installable in this image, so the oracle's C port stands in for the reference"},
           "e2e": {"value": round(value, 5), "unit": "Gcells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "integrals": vals}
    print(json.dumps(out))
