"""Committed golden fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py): seeded inputs and frozen oracle
outputs.  CPU: the live oracle still reproduces them (guards oracle drift) and the host-compiled evaluators match them.
GPU: the CUDA path, through the C ABI, matches them."""
import os

import numpy as np
import pytest

from helpers import State, run_ops, ocb, mirror
from cases import build_cases, check_case
from oracle.grid import OGrid, OField
from oracle import filters as OF, bpe as OB

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SWEEPS = [(False, "scalar"), (False, "tuple"), (True, "scalar"), (True, "tuple")]


def load(name):
    return dict(np.load(os.path.join(GOLD, name)))


def sweep_state(stretched, closure, device):
    fx = load(f"sweep_{'stretched' if stretched else 'regular'}_{closure}.npz")
    st = State(size=tuple(int(n) for n in fx["size"]), stretched=stretched, device=device)
    for name, f in (("u", st.ou), ("v", st.ov), ("w", st.ow), ("b", st.ob), ("p", st.op), ("nu", st.onu)):
        np.testing.assert_array_equal(f.data, fx["in_" + name], err_msg=f"seeded input {name} changed")
    return fx, st, build_cases(st, closure, mean="profile")


@pytest.mark.parametrize("stretched,closure", SWEEPS)
def test_oracle_and_host_evaluators_reproduce_the_fixtures(stretched, closure):
    fx, st, cases = sweep_state(stretched, closure, "cpu")
    vals, _ = run_ops([c[1] for c in cases], st.pg, backend="host")
    for (name, op, want, loc), got in zip(cases, vals):
        gold = fx["out_" + name]
        np.testing.assert_array_equal(want, gold, err_msg=f"oracle output {name} drifted from the committed fixture")
        check_case(st, name, got, gold)


def test_oracle_filters_and_bpe_reproduce_the_fixtures():
    fx = load("filters.npz")
    og = OGrid((10, 9, 8), x=(0, 1), y=(0, 1), z=(0, 1), halo=(2, 2, 2), topology=("P", "B", "B"))
    f = OField(og, "ccf", data=fx["in"].copy())
    np.testing.assert_array_equal(OF.box_filter(f, (1, 2, 3), 5, "shrink"), fx["box_N5_shrink"])
    np.testing.assert_array_equal(OF.gaussian_filter(f, (1, 3), 0.15, None, "shrink"), fx["gauss_s015_inferred_shrink"])
    fb = load("bpe.npz")
    og = OGrid((8, 6, 10), x=(0, 1), y=(0, 1), z=(-1, 0), topology=("P", "P", "B"))
    b = OField(og, "ccc", data=fb["in_b"].copy())
    r = OB.reference_height(b, "ThreeDimensionalSort")
    np.testing.assert_array_equal(r["zstar"], fb["ThreeDimensionalSort_zstar"])
    np.testing.assert_array_equal(r["perm"], fb["ThreeDimensionalSort_perm"])


@pytest.mark.gpu
@pytest.mark.parametrize("stretched,closure", SWEEPS)
def test_cuda_sweep_matches_the_fixtures(stretched, closure):
    fx, st, cases = sweep_state(stretched, closure, "cuda")
    vals, _ = run_ops([c[1] for c in cases], st.pg, backend="cuda", fused=True)
    for (name, op, want, loc), got in zip(cases, vals):
        check_case(st, name, got, fx["out_" + name])


@pytest.mark.gpu
def test_cuda_filters_and_bpe_match_the_fixtures():
    import torch
    SF, BPE = ocb.SpatialFilters, ocb.BackgroundPotentialEnergyEquation
    fx = load("filters.npz")
    og = OGrid((10, 9, 8), x=(0, 1), y=(0, 1), z=(0, 1), halo=(2, 2, 2), topology=("P", "B", "B"))
    pg = ocb.RectilinearGrid((10, 9, 8), x=(0, 1), y=(0, 1), z=(0, 1), halo=(2, 2, 2), topology=(ocb.Periodic, ocb.Bounded, ocb.Bounded), device="cuda")
    pf = mirror(OField(og, "ccf", data=fx["in"].copy()), pg)
    run = lambda op: ocb.Field(op).compute().interior_ijk()
    np.testing.assert_allclose(run(SF.BoxFilter(pf, dims=(1, 2, 3), N=5)), fx["box_N5_shrink"], rtol=1e-13, atol=1e-15)
    np.testing.assert_allclose(run(SF.BoxFilter(pf, dims=(2, 3), N=3, boundary="edge")), fx["box_N3_edge"], rtol=1e-13, atol=1e-15)
    np.testing.assert_allclose(run(SF.GaussianFilter(pf, dims=(1, 2, 3), sigma=0.2, N=(5, 7, 5), boundary=dict(left=-1.0, right=2.0))),
                               fx["gauss_s02_N575_const"], rtol=1e-13, atol=1e-15)
    np.testing.assert_allclose(run(SF.GaussianFilter(pf, dims=(1, 3), sigma=0.15)), fx["gauss_s015_inferred_shrink"], rtol=1e-13, atol=1e-15)
    fb = load("bpe.npz")
    og = OGrid((8, 6, 10), x=(0, 1), y=(0, 1), z=(-1, 0), topology=("P", "P", "B"))
    pg = ocb.RectilinearGrid((8, 6, 10), x=(0, 1), y=(0, 1), z=(-1, 0), device="cuda")
    pb = mirror(OField(og, "ccc", data=fb["in_b"].copy()), pg)
    for method in ("ThreeDimensionalSort", "HeavisideIntegral"):
        z = BPE.reference_height(pb, method=getattr(BPE, method)())
        torch.cuda.synchronize()
        np.testing.assert_array_equal(z.sorted_buoyancy.cpu().numpy(), fb[method + "_b_sorted"])
        np.testing.assert_array_equal(z.permutation.cpu().numpy() - 1, fb[method + "_perm"])
        np.testing.assert_allclose(z.interior_ijk(), fb[method + "_zstar"], rtol=0, atol=1e-12)
        np.testing.assert_allclose(z.reference_potential.interior_ijk(), fb[method + "_psi_delta"], rtol=1e-9, atol=1e-17)
