"""The C/OpenMP twin of the oracle (oracle/c/oracle_budget.c, the timed CPU baseline) agrees with the NumPy oracle,
which in turn is pinned to the reference's known-answer tests (test_oracle_golden.py)."""
import numpy as np
import pytest

from helpers import State, K
from oracle import cbudget


@pytest.mark.parametrize("size", [(12, 10, 8), (17, 9, 5)], ids=str)
def test_c_twin_matches_numpy_oracle(size):
    st = State(size=size)
    oc, _ = st.closures()["scalar"]
    cs = cbudget.from_oracle_fields(st.og, st.ou, st.ov, st.ow, st.ob, st.op, oc.nu, oc.kappa)
    og, f = st.og, st.ofields()
    want = {"KE": K.evaluate(K.kinetic_energy_ccc, og, "ccc", st.ou, st.ov, st.ow),
            "ADV": K.evaluate(K.ke_advection_ccc, og, "ccc", f),
            "PR": K.evaluate(K.ke_pressure_ccc, og, "ccc", f, st.op),
            "WB": K.evaluate(K.ke_buoyancy_ccc, og, "ccc", f, (0, 0, 1), st.ob),
            "EPS": K.evaluate(K.viscous_dissipation_rate_ccc, og, "ccc", None, f, {"closure": oc}),
            "CHI": K.evaluate(K.tracer_variance_dissipation_rate_ccc, og, "ccc", oc, None, None, st.ob),
            "CDIFF": K.evaluate(K.c_div_q_c, og, "ccc", oc, None, None, st.ob)}
    for name, w in want.items():
        got = cs.field(name)
        scale = np.max(np.abs(w))
        assert np.max(np.abs(got - w)) <= 1e-13 * scale, name
        integ = cs.integral(name)
        ref = K.integral(w, og, "ccc")
        i, j, k = og.indices("ccc")
        assert abs(integ - ref) <= 1e-13 * float(np.sum(np.abs(w) * og.V("ccc", i, j, k))), name
