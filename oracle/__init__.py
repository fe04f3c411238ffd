"""CPU oracle for the Oceanostics diagnostic hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is imported, linked or
executed by the product package (``oceanostics.jl_b200``).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may use it, and there only as the checker / the timed CPU
baseline -- never as the thing shipped.

What it is: a plain NumPy restatement (plus a C/OpenMP twin in ``oracle/c``) of
the reference's per-cell kernel functions, written operator-by-operator in the
same evaluation order as the Julia source, *without* any of the sharing or
re-association the CUDA kernels use.

Parity status ("pinned" vs "unpinned"):
  * The arithmetic of the reference's kernels lives in Oceananigans.jl
    (Project.toml:9,16 -- versions 0.102 .. 0.110, no exact pin; Manifest.toml
    is git-ignored), which is NOT vendored under /root/reference, and Julia is
    not installed in the build container.  The Oceananigans operators
    (delta / interpolation / derivative / areas / volumes, isotropic viscous and
    diffusive fluxes, centred second-order momentum advection, halo-fill
    semantics) are therefore restated from the published Oceananigans source
    (v0.10x series).
  * The oracle is pinned against every known-answer test and golden vector the
    reference's own test-suite / doctests hold for this path (see
    tests/test_oracle_golden.py; SURVEY.md section 8c lists them).
  * Element-wise outputs on *random* fields have no stored golden file in the
    reference (its tests only assert ``isa Field`` there), so for those the
    parity is "unpinned": it rests on the operator restatement being right.
"""
from .grid import OGrid, OField, ConstantField, ZeroField  # noqa: F401
