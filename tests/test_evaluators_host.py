"""CPU-container check of the library's per-cell evaluators (csrc/ocb_diag.cuh) against the NumPy oracle.

The evaluators are the exact templates the CUDA sweep kernel instantiates; tests/hostcheck compiles them for
the host so a wrong stencil is caught here, without a GPU.  The parity tests proper (through the C ABI, on
the device) are in test_sweep_gpu.py.  Tolerance: rtol 1e-12 in Float64 (BASELINE.json north_star), relative
to the field's magnitude."""
import numpy as np
import pytest

from helpers import State, run_ops, assert_close, K
from cases import build_cases, integral_scale, check_case


@pytest.mark.parametrize("stretched", [False, True], ids=["regular", "stretched"])
@pytest.mark.parametrize("closure", ["scalar", "smagorinsky_like", "tuple"])
def test_every_diagnostic_matches_oracle(stretched, closure):
    st = State(size=(10, 8, 7), stretched=stretched)
    cases = build_cases(st, closure, mean="profile")
    vals, ints = run_ops([c[1] for c in cases], st.pg, integrals=True, backend="host")
    for (name, op, want, loc), got, integ in zip(cases, vals, ints):
        finite = np.isfinite(want)
        check_case(st, name, got, want)
        if finite.all() and not name.startswith("RichardsonNumber"):
            assert abs(integ - K.integral(want, st.og, loc)) <= 1e-12 * integral_scale(st.og, want, loc), name


@pytest.mark.parametrize("mean", ["number", "none"])
def test_mean_flow_operand_kinds(mean):
    st = State(size=(8, 8, 6))
    cases = build_cases(st, "scalar", mean=mean)
    vals, _ = run_ops([c[1] for c in cases], st.pg, backend="host")
    for (name, op, want, loc), got in zip(cases, vals):
        check_case(st, name, got, want)


def test_float32_within_1e5_relative_l2():
    st = State(size=(10, 8, 7), stretched=True, dtype=np.float32)
    cases = build_cases(st, "tuple")
    vals, _ = run_ops([c[1] for c in cases], st.pg, backend="host")
    for (name, op, want32, loc), got in zip(cases, vals):
        if name.startswith(("RichardsonNumber", "QVelocity")):
            continue   # ratios of / differences between O(ε) quantities: not meaningful in Float32 on noise
        finite = np.isfinite(want32)
        num = np.linalg.norm((got.astype(np.float64) - want32.astype(np.float64))[finite])
        den = np.linalg.norm(want32.astype(np.float64)[finite])
        assert num <= 1e-5 * den, (name, num / den)
