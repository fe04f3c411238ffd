"""Host logic of the kinetic-energy forcing term: `Forcing(func, field_dependencies, parameters)`
functors are materialised by `operations.ForcingField` at every stored point of the forced field -- the values the oracle's
ContinuousForcing restatement gives there -- and `KineticEnergyForcing` wires them into the sweep operands (aux 0..2).
Runs without a GPU: the materialisation is array plumbing, only the sweep itself needs the CUDA library."""
import numpy as np
import pytest
import torch

from helpers import State, K, ocb
from oceanostics_b200.operations import ForcingField, forcing_operand
from oceanostics_b200.fields import ZeroField, ConstantField


def test_forcing_field_equals_the_oracle_functor_at_every_stored_point():
    st = State(size=(12, 10, 8), device="cpu")
    og = st.og
    relax = lambda x, y, z, t, q, p: -(p["r"] + 0.25 * np.sin(2 * np.pi * x) * np.cos(2 * np.pi * y) + z + t) * q
    trelax = lambda x, y, z, t, q, p: -(p["r"] + 0.25 * torch.sin(2 * torch.pi * x) * torch.cos(2 * torch.pi * y) + z + t) * q
    m = st.model(closure=st.closures()["scalar"][1])
    m.clock.time = 0.75
    clock = type("Clock", (), {"time": 0.75})()
    for name, loc, src, osrc in (("u", "fcc", st.u, st.ou), ("v", "cfc", st.v, st.ov), ("w", "ccf", st.w, st.ow)):
        ff = ForcingField(ocb.Forcing(trelax, field_dependencies=(name,), parameters={"r": 0.5}), src, m).compute()
        F = K.ContinuousForcing(relax, loc, (name,), {"r": 0.5})
        H = og.H
        ii, jj, kk = (np.arange(1 - H[d], osrc.data.shape[d] - H[d] + 1) for d in range(3))
        want = F(ii[:, None, None], jj[None, :, None], kk[None, None, :], og, clock, st.ofields())
        got = ff.data.numpy().transpose(2, 1, 0)
        assert got.shape == want.shape
        np.testing.assert_allclose(got, want, rtol=1e-14, atol=1e-16)
        assert ff.location == src.location and ff.is_computed


def test_forcing_operands_and_constructor_wiring():
    st = State(size=(8, 6, 4), device="cpu")
    m = st.model(closure=st.closures()["scalar"][1])
    assert isinstance(forcing_operand(None, st.u, m), ZeroField)
    assert isinstance(forcing_operand(0.25, st.u, m), ConstantField)
    assert forcing_operand(st.u, st.u, m) is st.u
    with pytest.raises(ValueError):
        forcing_operand(st.v, st.u, m)                      # a forcing Field lives where the forced field does
    m.forcing.u = ocb.Forcing(lambda x, y, z, t, u: -u, field_dependencies="u")
    m.forcing.w = 0.5
    op = ocb.KineticEnergyEquation.KineticEnergyForcing(m)
    assert op.name == "KineticEnergyForcing" and op.location == ("c", "c", "c")
    aux = op.operands.aux
    assert isinstance(aux[0], ForcingField) and isinstance(aux[1], ZeroField) and isinstance(aux[2], ConstantField)
    with pytest.raises(ValueError):
        ForcingField(ocb.Forcing(lambda x, y, z, t, v: -v, field_dependencies="v"), st.u, m).compute()
    with pytest.raises(ValueError):
        ocb.NonhydrostaticModel(st.pg, forcing={"q": 1.0})
    with pytest.raises(TypeError):
        ocb.KineticEnergyEquation.KineticEnergyForcing(ocb.HydrostaticFreeSurfaceModel(st.pg))
