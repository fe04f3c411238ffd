"""Host-side logic of the SpatialFilters constructors (no GPU): validation rules and error messages of
src/SpatialFilters/SpatialFilters.jl:234-271 and gaussian_filter.jl:488-505, pass construction, width inference."""
import numpy as np
import pytest

from helpers import ocb

SF = ocb.SpatialFilters


def grid(topology=(ocb.Periodic, ocb.Periodic, ocb.Bounded), size=(8, 8, 8)):
    return ocb.RectilinearGrid(size, x=(0, 1), y=(0, 1), z=(0, 1), topology=topology, device="cpu")


def test_validation_errors():
    """test/test_spatial_filters.jl validation block."""
    c = ocb.CenterField(grid())
    for bad_dims in [(), (4,), (1, 1), (0,), "x", (1.0,)]:
        with pytest.raises(ValueError):
            SF.BoxFilter(c, dims=bad_dims, N=3)
    for bad_N in [2, 4, 1, 3.0, True]:
        with pytest.raises(ValueError):
            SF.BoxFilter(c, dims=(1,), N=bad_N)
    for bad_sigma in [0, -1.0, None, "a"]:
        with pytest.raises(ValueError):
            SF.GaussianFilter(c, dims=(1,), sigma=bad_sigma)
    with pytest.raises(ValueError):
        SF.BoxFilter(c, dims=(3,), N=3, boundary="wrap")
    with pytest.raises(ValueError):
        SF.BoxFilter(c, dims=(3,), N=3, boundary=dict(left=1.0))
    with pytest.raises(ValueError):
        SF.BoxFilter(c, dims=(1, 3), N=3, boundary=("shrink",))
    with pytest.raises(ValueError):            # periodic width limit 2*Nd+1 (SpatialFilters.jl:248-257)
        SF.BoxFilter(c, dims=(1,), N=19)
    SF.BoxFilter(c, dims=(1,), N=17)
    for bad in [(5,), (5, 7, 9), (3, 4), (5, 1)]:
        with pytest.raises(ValueError):
            SF.GaussianFilter(c, dims=(1, 3), sigma=0.2, N=bad)


def test_policies_extents_and_inferred_widths():
    g = grid()
    w = ocb.ZFaceField(g)
    op = SF.GaussianFilter(w, dims=(3, 1), sigma=1.5 / 8, boundary=("edge", "shrink"))
    assert op.sorted_dims == (1, 3)
    assert op.policies[0][0] == "periodic" and op.policies[0][-1] == 8            # periodic wins over the user's spec
    assert op.policies[1][0] == "edge" and op.policies[1][-1] == 9                # Face x Bounded: N + 1 points
    assert op.widths == (3, 3)                                                   # ceil(2σ/Δ) = ceil(3) = 3
    passes, _ = op.build_passes()
    assert [p.dim for p in passes] == [1, 3] and passes[1].extent == 9
    assert passes[0].kind == ocb._lib.OCB_FILTER_GAUSSIAN and abs(passes[0].weights[3] - 1.0) < 1e-15
    assert op.location == ("c", "c", "f")


def test_operator_form_validates_eagerly():
    with pytest.raises(ValueError):
        SF.BoxFilter(dims=(1,), N=4)
    with pytest.raises(ValueError):
        SF.GaussianFilter(dims=(5,), sigma=0.1)
    f = SF.GaussianFilter(dims=(1, 2), sigma=0.1)
    assert callable(f)
