"""Host check of the identity behind `ocb_budget_sums_kernel` (csrc/ocb_budget_sums.cuh): the seven budget integrals
over (all x) x (a y window) x (a z window) written as weighted sums over staggered items -- faces, ffc corners, fcf / cff
edges, cells -- with class weights that depend on the row and the plane only, the advection term summed by parts with its
explicit window-edge terms.  The NumPy transcription below uses the kernel's GEN formulas (its FAST form is the same
expression with every weight equal to one) and must reproduce the oracle's per-cell evaluation summed over the window
(src/KineticEnergyEquation.jl:37-39, 147-152, 304-309, 362-367, 427-453; src/TracerVarianceEquation.jl:93-98, 146-161)."""
import numpy as np
import pytest

from helpers import State, K


def class_sums(st, nu, kappa, jlo, jhi, klo, khi):
    og = st.og
    Nx, Ny, Nz = og.N
    H = og.H[0]
    rdx, rdy, rdz = Nx / 1.0, Ny / 1.0, Nz / 1.0
    I = np.arange(1, Nx + 1)[:, None, None]
    J = np.arange(jlo, jhi + 2)[None, :, None]                  # rows jlo .. jhi + 1 carry items
    N = np.arange(klo, khi + 3)[None, None, :]                  # steps klo .. khi + 2

    def at(f, di=0, dj=0, dk=0):
        return f.data[I + di + H - 1, J + dj + H - 1, N + dk + H - 1]

    u, v, w, b, p = st.ou, st.ov, st.ow, st.ob, st.op
    inw = lambda a, lo, hi: ((a >= lo) & (a <= hi)).astype(np.float64)
    al, be, an = inw(J, jlo, jhi), inw(J - 1, jlo, jhi), inw(J + 1, jlo, jhi)
    g0, g1, g2 = inw(N, klo, khi), inw(N - 1, klo, khi), inw(N - 2, klo, khi)
    hy, hyn, hz, hzp = 0.5 * (al + be), 0.5 * (al + an), 0.5 * (g0 + g1), 0.5 * (g1 + g2)
    uc, us, up_, ue = at(u), at(u, dj=-1), at(u, dk=-1), at(u, di=1)
    vc, vw, vp, vn, vs = at(v), at(v, di=-1), at(v, dk=-1), at(v, dj=1), at(v, dj=-1)
    wc, ww, ws, wp = at(w), at(w, di=-1), at(w, dj=-1), at(w, dk=-1)
    bc, bw, bs, bp, be_, bn = at(b), at(b, di=-1), at(b, dj=-1), at(b, dk=-1), at(b, di=1), at(b, dj=1)
    pc, pw, ps, pp = at(p), at(p, di=-1), at(p, dj=-1), at(p, dk=-1)
    duy, dvx, duz, dwx, dvz, dwy = uc - us, vc - vw, uc - up_, wc - ww, vc - vp, wc - ws
    ax, ay, dwz = ue - uc, vn - vc, wc - wp
    S12, S13, S23 = dvx * rdx + duy * rdy, dwx * rdx + duz * rdz, dwy * rdy + dvz * rdz
    P12, P13, P23 = (vw + vc) * (us + uc), (ww + wc) * (up_ + uc), (ws + wc) * (vp + vc)
    Uu, Vv, Ww, VvS = (uc + ue) ** 2, (vc + vn) ** 2, (wp + wc) ** 2, (vs + vc) ** 2
    S = np.sum
    aEe = S(hy * g0 * S12 ** 2 + al * hz * S13 ** 2 + hy * hz * S23 ** 2)
    aEx, aEy, aEz = S(al * g0 * ax ** 2), S(al * g0 * ay ** 2), S(al * g1 * dwz ** 2)
    aAe = S(g0 * P12 * (rdx * hy * dvx + rdy * (al * uc - be * us)) + al * P13 * (rdx * hz * dwx + rdz * (g0 * uc - g1 * up_)) +
            P23 * (rdy * hz * (al * wc - be * ws) + rdz * hy * (g0 * vc - g1 * vp)))
    aAx = S(al * g0 * Uu * ax)
    aAy = S(g0 * Vv * (hyn * vn - hy * vc)) + S((J == jlo) * g0 * VvS * 0.5 * vc)
    aAz = S(al * Ww * (hz * wc - hzp * wp))
    aK = S(al * g0 * uc ** 2 + hy * g0 * vc ** 2 + al * hz * wc ** 2)
    aPx, aPy, aPz = S(al * g0 * uc * (pc - pw)), S(hy * g0 * vc * (pc - ps)), S(al * hz * wc * (pc - pp))
    aW = S(al * hz * wc * (bp + bc))
    dbx, dby, dbz = bc - bw, bc - bs, bc - bp
    aCx, aCy, aCz = S(al * g0 * dbx ** 2), S(hy * g0 * dby ** 2), S(al * hz * dbz ** 2)
    # Laplacian (sign of div q) of the cell below the face of step n: in-plane part of plane n - 1, z faces n - 1 and n
    bpw, bpe, bps, bpn, bpp = at(b, di=-1, dk=-1), at(b, di=1, dk=-1), at(b, dj=-1, dk=-1), at(b, dj=1, dk=-1), at(b, dk=-2)
    lap = rdx ** 2 * ((bp - bpw) - (bpe - bp)) + rdy ** 2 * ((bp - bps) - (bpn - bp)) + rdz ** 2 * ((bp - bpp) - dbz)
    aD = S(al * g1 * bp * lap)
    V = 1.0 / (rdx * rdy * rdz)
    return [V * 0.5 * aK, V * -0.25 * (aAe + rdx * aAx + rdy * aAy + rdz * aAz), V * (rdx * aPx + rdy * aPy + rdz * aPz), V * 0.5 * aW,
            V * nu * (aEe + 2.0 * (rdx ** 2 * aEx + rdy ** 2 * aEy + rdz ** 2 * aEz)),
            V * 2.0 * kappa * (rdx ** 2 * aCx + rdy ** 2 * aCy + rdz ** 2 * aCz), V * 2.0 * kappa * aD]


@pytest.mark.parametrize("window", [None, (3, 9, 1, 10), (1, 1, 2, 2), (12, 12, 10, 10), (1, 12, 4, 7), (2, 11, 1, 1)], ids=str)
def test_class_sums_reproduce_the_windowed_cell_integrals(window):
    st = State(size=(16, 12, 10))
    og, f = st.og, st.ofields()
    nu, kappa = 1e-6, 1e-7
    oc = K.Closure(nu=nu, kappa=kappa)
    ev = K.evaluate
    cells = [ev(K.kinetic_energy_ccc, og, "ccc", st.ou, st.ov, st.ow), ev(K.ke_advection_ccc, og, "ccc", f),
             ev(K.ke_pressure_ccc, og, "ccc", f, st.op), ev(K.ke_buoyancy_ccc, og, "ccc", f, (0, 0, 1), st.ob),
             ev(K.viscous_dissipation_rate_ccc, og, "ccc", None, f, {"closure": oc}),
             ev(K.tracer_variance_dissipation_rate_ccc, og, "ccc", oc, None, None, st.ob), ev(K.c_div_q_c, og, "ccc", oc, None, None, st.ob)]
    jlo, jhi, klo, khi = window or (1, 12, 1, 10)
    got = class_sums(st, nu, kappa, jlo, jhi, klo, khi)
    V = 1.0 / (16 * 12 * 10)
    for name, c, g in zip("K ADV PR WB EPS CHI CDF".split(), cells, got):
        win = c[:, jlo - 1:jhi, klo - 1:khi]
        want, scale = float(np.sum(win)) * V, float(np.sum(np.abs(win))) * V
        assert abs(g - want) <= 1e-12 * scale, (name, g, want, scale)
