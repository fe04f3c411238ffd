"""CPU arm of bench.py: the oracle's C/OpenMP port (oracle/c/oracle_budget.c) timed on the host cores.

The reference itself is Julia + Oceananigans and cannot run in this image (no Julia), so the
"reference" arm is the declared stand-in the task contract allows: the per-cell, unshared, one-sweep-per-diagnostic
restatement of the reference's kernel functions, on ALL host threads (torchrun's OMP_NUM_THREADS=1 is overridden), each
step a bounded sample of the same workload.
"""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
from bench_config import METRIC, UNIT, DEFAULT_SIZE, workload_config  # noqa: E402


def _host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def _all_threads():
    """Every core this process may run on, whatever OMP_NUM_THREADS says (torchrun sets it to 1 for its workers)."""
    from oracle import cbudget
    n = _host_cores()
    cbudget.lib().ob_set_num_threads(n)
    return cbudget.lib().ob_num_threads()


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

    def noise(a, nz, amp):          # plane by plane: bounded temporaries at 512^3
        for k in range(nz):
            a[H + k, H:H + n, H:H + n] += amp * rng.standard_normal((n, n), dtype=np.float64)

    def fill(a, nz, wall_normal=False):
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
    u[H:H + n, H:H + n, H:H + n] = np.sin(two_pi * xf)[None, None, :] * np.cos(two_pi * xc)[None, :, None]
    v[H:H + n, H:H + n, H:H + n] = -np.cos(two_pi * xc)[None, None, :] * np.sin(two_pi * xf)[None, :, None]
    b[H:H + n, H:H + n, H:H + n] = 1e-5 * zc[:, None, None]
    noise(u, n, 0.1); noise(v, n, 0.1); noise(w, n + 1, 0.1); noise(b, n, 1e-6); noise(p, n, 0.1)
    fill(u, n); fill(v, n); fill(w, n + 1, wall_normal=True); fill(b, n); fill(p, n)
    d = (1.0 / n, 1.0 / n, 1.0 / n)
    return cbudget.BudgetState((n, n, n), H, d, 1e-6, 1e-7, u, v, w, b, p)


def _one_step(state):
    """One step of the CPU arm: the 7 budget integrals, one sweep per integral (the reference does not fuse)."""
    from oracle import cbudget
    t0 = time.perf_counter()
    vals = [state.integral(d) for d in cbudget.DIAGS]
    return time.perf_counter() - t0, vals


def cpu_baseline(args):
    """bench.py's `cpu_baseline` object: the port timed on a bounded sample (rank 0, N=1 only)."""
    import __graft_entry__ as ge
    ge.build_oracle()
    cores = _all_threads()
    n = int(args.cpu_size or 256)
    st = _host_state(n)
    _one_step(st)            # warm-up (page faults, OpenMP team start)
    best = min(_one_step(st)[0] for _ in range(3))
    return {"value": round(n ** 3 / best / 1e9, 5), "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{n}^3 Float64 sample of the same workload; 7 budget integrals, one OpenMP sweep per integral "
                      f"(the reference does not fuse), best of 3 after one warm-up; {best:.3f} s per step"}


def run_reference(args):
    """`bench.py --impl reference`: same metric / unit / config line, timed on the host cores (rank 0 only), exactly
    `--warmup` + `--steps` steps, each step the 7 integrals on a bounded sample cube of the workload."""
    if int(os.environ.get("RANK", 0)) != 0:
        return
    import __graft_entry__ as ge
    ge.build_oracle()
    cores = _all_threads()
    # sample edge: a step of ~1-3 s on the box's cores, so that W + K steps end within a few minutes
    n = int(args.cpu_size or 512)
    if args.cpu_size is None and (args.steps + max(args.warmup, 3)) > 40:
        n = 384
    st = _host_state(n)
    steps, warmup = args.steps, max(args.warmup, 3)
    for _ in range(warmup):
        _one_step(st)
    t0 = time.perf_counter()
    for _ in range(steps):
        _, vals = _one_step(st)
    dt = (time.perf_counter() - t0) / steps
    value = n ** 3 / dt / 1e9
    size = tuple(args.size) if args.size else DEFAULT_SIZE
    world = int(os.environ.get("WORLD_SIZE", args.gpus or 1))
    sample = (f"each step = the 7 budget integrals on a {n}^3 Float64 sample cube of the workload (same fields, closure, "
              f"topology), one OpenMP sweep per integral as the reference's one compute! per diagnostic, {cores} threads")
    out = {"impl": "reference", "metric": METRIC, "value": round(value, 5), "unit": UNIT, "n_gpus": world, "steps": steps,
           "warmup": warmup, "ms_per_step": round(dt * 1e3, 3), "higher_is_better": True, "scaling": "strong" if world > 1 else "weak",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": workload_config(size, world, args.fields),
           "cpu_baseline": {"value": round(value, 5), "unit": UNIT, "cores": cores, "kind": "port",
                            "sample": sample + "; Julia/Oceananigans are not installable in this image, so the oracle's C port "
                                               "stands in for the reference (Gcells/s is size-independent for this streaming sweep)"},
           "e2e": {"value": round(value, 5), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "integrals_of_the_sample": vals}
    print(json.dumps(out))
