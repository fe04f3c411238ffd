#!/usr/bin/env python
"""Generate the committed golden fixtures: seeded inputs and the NumPy oracle's outputs for every K1 diagnostic
(regular and stretched grid, scalar and eddy-viscosity closures), the separable filters and the sorted reference
height.  The reference itself (Julia + Oceananigans) cannot run in the build image, so these are outputs of the
oracle -- which is pinned to the reference's own known-answer tests in tests/test_oracle_golden.py -- frozen so that
later changes of either the oracle or the CUDA path are caught.  Run from the repo root:

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from helpers import State  # noqa: E402
from cases import build_cases  # noqa: E402
from oracle.grid import OGrid, OField  # noqa: E402
from oracle import filters as OF, bpe as OB  # noqa: E402

SIZE = (12, 10, 8)


def sweep_fixture(stretched, closure):
    st = State(size=SIZE, stretched=stretched)
    cases = build_cases(st, closure, mean="profile")
    out = {"size": np.array(SIZE), "stretched": np.array(stretched)}
    for name, f in (("u", st.ou), ("v", st.ov), ("w", st.ow), ("b", st.ob), ("p", st.op), ("nu", st.onu)):
        out["in_" + name] = f.data          # parent arrays [i, j, k] with halos
    for name, op, want, loc in cases:
        out["out_" + name] = want
    return out


def filter_fixture():
    rng = np.random.default_rng(7)
    og = OGrid((10, 9, 8), x=(0, 1), y=(0, 1), z=(0, 1), halo=(2, 2, 2), topology=("P", "B", "B"))
    f = OField(og, "ccf")
    f.interior[...] = rng.standard_normal(f.interior.shape)
    f.fill_halo(impenetrable=False)
    return {"in": f.data,
            "box_N5_shrink": OF.box_filter(f, (1, 2, 3), 5, "shrink"),
            "box_N3_edge": OF.box_filter(f, (2, 3), 3, "edge"),
            "gauss_s02_N575_const": OF.gaussian_filter(f, (1, 2, 3), 0.2, (5, 7, 5), dict(left=-1.0, right=2.0)),
            "gauss_s015_inferred_shrink": OF.gaussian_filter(f, (1, 3), 0.15, None, "shrink")}


def bpe_fixture():
    rng = np.random.default_rng(11)
    og = OGrid((8, 6, 10), x=(0, 1), y=(0, 1), z=(-1, 0), topology=("P", "P", "B"))
    b = OField(og, "ccc")
    b.interior[...] = 1e-5 * (-1 + (np.arange(10) + 0.5) / 10) + 1e-6 * rng.standard_normal(b.interior.shape)
    b.interior[2:4, :, 5] = b.interior[2, 0, 5]       # a few ties
    b.fill_halo()
    out = {"in_b": b.data}
    for method in ("ThreeDimensionalSort", "HeavisideIntegral"):
        r = OB.reference_height(b, method)
        out[method + "_zstar"] = r["zstar"]
        out[method + "_psi_delta"] = r["psi_delta"]
        out[method + "_b_sorted"] = r["b_sorted"]
        out[method + "_perm"] = r["perm"]
    return out


def main():
    for stretched in (False, True):
        for closure in ("scalar", "tuple"):
            name = f"sweep_{'stretched' if stretched else 'regular'}_{closure}.npz"
            np.savez_compressed(os.path.join(HERE, name), **sweep_fixture(stretched, closure))
    np.savez_compressed(os.path.join(HERE, "filters.npz"), **filter_fixture())
    np.savez_compressed(os.path.join(HERE, "bpe.npz"), **bpe_fixture())
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
