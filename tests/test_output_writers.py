"""Output side of the path: schedules, `indices=` windows, time-averaged windows and the
containers on the CPU with plain Fields (no kernel runs); on the GPU, computed outputs -- windows equal the matching slices
of the full fields (test/test_spatial_filters.jl:207-222), siblings share one sweep, and every output step sees the evolved
state (test/test_subfilter_kinetic_energy_equation.jl:87)."""
import os

import numpy as np
import pytest
import torch

from helpers import State, ocb, K


def _cpu_model():
    g = ocb.RectilinearGrid((6, 5, 4), extent=(1, 1, 1), device="cpu")
    return ocb.NonhydrostaticModel(g)


def test_schedules():
    clock = _cpu_model().clock
    it = ocb.IterationInterval(3)
    hits = []
    for n in range(7):
        clock.iteration = n
        hits.append(it(clock))
    assert hits == [True, False, False, True, False, False, True]
    ti = ocb.TimeInterval(0.5)
    hits = []
    for t in (0.0, 0.2, 0.4, 0.5, 0.7, 1.0, 1.1):
        clock.time = t
        hits.append(ti(clock))
    assert hits == [True, False, False, True, False, True, False]
    with pytest.raises(ValueError):
        ocb.AveragedTimeInterval(1.0, window=2.0)


def test_windows_time_average_and_containers(tmp_path):
    m = _cpu_model()
    b = m.tracers.b
    b.interior.copy_(torch.arange(b.interior.numel(), dtype=b.interior.dtype).reshape(b.interior.shape))
    w = ocb.NetCDFWriter(m, {"b": b}, filename="snap.nc", schedule=ocb.IterationInterval(1), indices=(None, None, 4), dir=str(tmp_path))
    j = ocb.JLD2Writer(m, {"b": b}, filename="snap", schedule=ocb.IterationInterval(2), indices=(None, (2, 3), None), dir=str(tmp_path))
    avg = ocb.NetCDFWriter(m, {"b": b}, filename="avg.nc", schedule=ocb.AveragedTimeInterval(1.0, window=0.5), dir=str(tmp_path))
    want_avg, acc, elapsed = [], 0.0, 0.0
    for n in range(9):
        m.clock.iteration, m.clock.time = n, 0.25 * n
        b.interior.add_(1.0)
        w(m); j(m); avg(m)
    from scipy.io import netcdf_file
    with netcdf_file(os.path.join(tmp_path, "snap.nc"), "r", mmap=False) as nc:
        v = nc.variables["b"][:]
        assert v.shape == (9, 1, 5, 6) and nc.variables["time"][:].tolist() == [0.25 * n for n in range(9)]
        first = np.arange(120, dtype=np.float64).reshape(4, 5, 6)[3:4] + 1.0
        np.testing.assert_array_equal(v[0], first)
        np.testing.assert_array_equal(v[8], first + 8.0)
    z = np.load(os.path.join(tmp_path, "snap.npz"))
    assert z["timeseries/b/0"].shape == (4, 2, 6) and float(z["timeseries/t/2"]) == 1.0
    np.testing.assert_array_equal(z["timeseries/b/1"], np.arange(120.).reshape(4, 5, 6)[:, 1:3, :] + 3.0)
    # windows (0.5, 1.0] and (1.5, 2.0]: steps at t = 0.75, 1.0 and 1.75, 2.0, each weighted by dt = 0.25
    with netcdf_file(os.path.join(tmp_path, "avg.nc"), "r", mmap=False) as nc:
        v, t = nc.variables["b"][:], nc.variables["time"][:]
        assert t.tolist() == [1.0, 2.0] and v.shape == (2, 4, 5, 6)
        base = np.arange(120.).reshape(4, 5, 6)
        np.testing.assert_allclose(v[0], base + 0.5 * (4 + 5))
        np.testing.assert_allclose(v[1], base + 0.5 * (8 + 9))


@pytest.mark.gpu
def test_computed_outputs_are_windowed_fused_and_follow_the_state(tmp_path):
    st = State(size=(32, 24, 16), device="cuda")
    m = st.model(closure=st.closures()["scalar"][1])
    KE, TV, SF = ocb.KineticEnergyEquation, ocb.TracerVarianceEquation, ocb.SpatialFilters
    outs = {"eps": KE.DissipationRate(m), "K": KE.KineticEnergy(m), "chi": TV.TracerVarianceDissipationRate(m, "b"),
            "intK": ocb.Integral(KE.KineticEnergy(m)), "bbar": SF.GaussianFilter(st.b, dims=(1, 2, 3), sigma=2.0 / 32, N=5)}
    Nz = 16
    top = ocb.NetCDFWriter(m, outs, filename="top.nc", schedule=ocb.IterationInterval(1), indices=(None, None, Nz), dir=str(tmp_path))
    assert len(top.groups) == 1 and len(top.groups[0].members) == 4          # eps, K, chi and the integral share ONE sweep
    full = {k: ocb.Field(v) for k, v in outs.items() if k != "intK"}
    for n in range(2):
        m.clock.iteration, m.clock.time = n, float(n)
        assert top(m)
        for f in full.values():
            f.compute(time=float(n))
        torch.cuda.synchronize()
        for name, f in full.items():
            got = top.records[name][n]
            assert got.shape == (1, 24, 32)
            np.testing.assert_allclose(got, f.interior[Nz - 1:Nz].cpu().numpy(), rtol=1e-13, atol=0)
        want = K.integral(K.evaluate(K.kinetic_energy_ccc, st.og, "ccc", st.ou, st.ov, st.ow), st.og, "ccc") * (1.0 if n == 0 else 2.25)
        assert abs(float(top.records["intK"][n]) - want) <= 1e-12 * abs(want)
        st.u.data.mul_(1.5); st.v.data.mul_(1.5); st.w.data.mul_(1.5)     # the flow evolves between outputs
    from scipy.io import netcdf_file
    with netcdf_file(os.path.join(tmp_path, "top.nc"), "r", mmap=False) as nc:
        assert nc.variables["eps"][:].shape == (2, 1, 24, 32) and nc.variables["intK"][:].shape == (2,)
        assert not np.array_equal(nc.variables["K"][0], nc.variables["K"][1])
    # time average over a window on the device
    avg = ocb.JLD2Writer(m, {"K": KE.KineticEnergy(m)}, filename="avg", schedule=ocb.AveragedTimeInterval(1.0, window=1.0), dir=str(tmp_path))
    vals = []
    for n in range(5):
        m.clock.time = 2.0 + 0.25 * n
        avg(m)
        if n > 0:
            vals.append(ocb.Field(KE.KineticEnergy(m)).compute().interior.cpu().numpy())
        st.u.data.mul_(1.1)
    z = np.load(os.path.join(tmp_path, "avg.npz"))
    np.testing.assert_allclose(z["timeseries/K/0"], sum(vals) / 4, rtol=1e-13)
