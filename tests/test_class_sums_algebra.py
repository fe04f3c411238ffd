"""Host check of the identity behind `ocb_budget_sums_kernel` (csrc/ocb_budget_sums.cuh): the seven budget integrals
over (all x) x (a y window) x (a z window) written as weighted sums over staggered items -- faces, ffc corners, fcf / cff
edges, cells -- with class weights that depend on the row and the plane only, the advection term summed by parts with its
explicit window-edge terms.  The NumPy transcription below uses the kernel's GEN formulas (its FAST form is the same
expression with every weight equal to one) and must reproduce the oracle's per-cell evaluation summed over the window
(src/KineticEnergyEquation.jl:37-39, 147-152, 304-309, 362-367, 427-453; src/TracerVarianceEquation.jl:93-98, 146-161)."""
import numpy as np
import pytest

from helpers import State, K


def class_sums(st, nu, kappa, jlo, jhi, klo, khi, ring=False, profiles=None):
    """ring: the kernel's ring mode -- rows jlo .. jhi only, every row weight one (the y window is a whole periodic extent or
    one slab of a ring).  profiles = (U, V, W) oracle reduced fields: also returns the TKE-budget class sums (shear production
    summed by faces, dissipation of u - U, turbulent kinetic energy)."""
    og = st.og
    Nx, Ny, Nz = og.N
    H = og.H[0]
    rdx, rdy, rdz = Nx / 1.0, Ny / 1.0, Nz / 1.0
    I = np.arange(1, Nx + 1)[:, None, None]
    J = np.arange(jlo, jhi + (1 if ring else 2))[None, :, None]  # rows jlo .. jhi + 1 carry items (ring mode: the next slab's / wrapped row counts them)
    N = np.arange(klo, khi + 3)[None, None, :]                  # steps klo .. khi + 2

    def at(f, di=0, dj=0, dk=0):
        return f.data[I + di + H - 1, J + dj + H - 1, N + dk + H - 1]

    u, v, w, b, p = st.ou, st.ov, st.ow, st.ob, st.op
    inw = lambda a, lo, hi: ((a >= lo) & (a <= hi)).astype(np.float64)
    al, be, an = inw(J, jlo, jhi), inw(J - 1, jlo, jhi), inw(J + 1, jlo, jhi)
    if ring:
        al, be, an = np.ones_like(al), np.ones_like(al), np.ones_like(al)
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
    aAy = S(g0 * Vv * (hyn * vn - hy * vc)) + (0.0 if ring else S((J == jlo) * g0 * VvS * 0.5 * vc))
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
    extra = []
    if profiles is not None:
        U, Vp, W = profiles
        prof = lambda f, dk=0: f.data[0, 0, N[0, 0, :] + dk + H - 1][None, None, :]
        Un, Un1, Vn, Vn1, Wn = prof(U), prof(U, -1), prof(Vp), prof(Vp, -1), prof(W)
        GUc, GUp = 0.5 * (prof(U, 1) - Un1), 0.5 * (Un - prof(U, -2))
        GVc, GVp = 0.5 * (prof(Vp, 1) - Vn1), 0.5 * (Vn - prof(Vp, -2))
        GWc, GWp = prof(W, 1) - Wn, Wn - prof(W, -1)
        ucq, vcq, wcq, upq, vpq = uc - Un, vc - Vn, wc - Wn, up_ - Un1, vp - Vn1
        S13q, S23q, dwzq = S13 - rdz * (Un - Un1), S23 - rdz * (Vn - Vn1), dwz - GWp
        aEeP = S(hy * g0 * S12 ** 2 + al * hz * S13q ** 2 + hy * hz * S23q ** 2)
        aEzP = S(al * g1 * dwzq ** 2)
        aKP = S(al * g0 * ucq ** 2 + hy * g0 * vcq ** 2 + al * hz * wcq ** 2)
        aSuv = S(al * ((ww - Wn) + wcq) * (g0 * GUc * ucq + g1 * GUp * upq)) + S((be * (ws - Wn) + al * wcq) * (g0 * GVc * vcq + g1 * GVp * vpq))
        aSw = S(al * wcq ** 2 * (g0 * GWc + g1 * GWp))
        extra = [V * -rdz * (0.25 * aSuv + 0.5 * aSw), V * nu * (aEeP + 2.0 * (rdx ** 2 * aEx + rdy ** 2 * aEy + rdz ** 2 * aEzP)), V * 0.5 * aKP]
    return extra + [V * 0.5 * aK, V * -0.25 * (aAe + rdx * aAx + rdy * aAy + rdz * aAz), V * (rdx * aPx + rdy * aPy + rdz * aPz), V * 0.5 * aW,
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


def _budget_cells(st, oc):
    og, f = st.og, st.ofields()
    ev = K.evaluate
    return [ev(K.kinetic_energy_ccc, og, "ccc", st.ou, st.ov, st.ow), ev(K.ke_advection_ccc, og, "ccc", f),
            ev(K.ke_pressure_ccc, og, "ccc", f, st.op), ev(K.ke_buoyancy_ccc, og, "ccc", f, (0, 0, 1), st.ob),
            ev(K.viscous_dissipation_rate_ccc, og, "ccc", None, f, {"closure": oc}),
            ev(K.tracer_variance_dissipation_rate_ccc, og, "ccc", oc, None, None, st.ob), ev(K.c_div_q_c, og, "ccc", oc, None, None, st.ob)]


def test_ring_mode_shares_add_up_to_the_global_integrals():
    """Ring mode of the kernel (csrc/ocb_budget_sums.cuh, `y_ring`): with every row weight one and no extra row, the whole
    periodic y extent gives the global integrals, and the shares of y slabs -- each reading ONE row beyond its slab -- add up to
    them, although a share is not the integral over its slab's cells (the advection term moves momentum across the cut)."""
    st = State(size=(16, 12, 10))
    nu, kappa = 1e-6, 1e-7
    cells = _budget_cells(st, K.Closure(nu=nu, kappa=kappa))
    V = 1.0 / (16 * 12 * 10)
    whole = class_sums(st, nu, kappa, 1, 12, 1, 10, ring=True)
    slabs = [class_sums(st, nu, kappa, lo, hi, 1, 10, ring=True) for lo, hi in ((1, 4), (5, 8), (9, 12))]
    for n, (name, c) in enumerate(zip("K ADV PR WB EPS CHI CDF".split(), cells)):
        want, scale = float(np.sum(c)) * V, float(np.sum(np.abs(c))) * V
        assert abs(whole[n] - want) <= 1e-12 * scale, (name, whole[n], want)
        assert abs(sum(s[n] for s in slabs) - want) <= 1e-12 * scale, name
    adv_cells_of_slab = float(np.sum(cells[1][:, 0:4, :])) * V
    assert abs(slabs[0][1] - adv_cells_of_slab) > 1e-9 * float(np.sum(np.abs(cells[1][:, 0:4, :]))) * V


@pytest.mark.parametrize("window", [None, (3, 9, 1, 10), (1, 1, 2, 2), (12, 12, 10, 10), (2, 11, 3, 8)], ids=str)
def test_tke_budget_class_sums_reproduce_the_cell_integrals(window):
    """The TKE-budget terms of the class-sum kernel about U(z), V(z), W(z): shear production summed over x / y / z faces, the
    dissipation of u - U, the turbulent kinetic energy -- against the oracle's per-cell kernels
    (src/TurbulentKineticEnergyEquation.jl:24-30, 224-240, 285-289; src/KineticEnergyEquation.jl:441-453) summed over the window."""
    from oracle.grid import average_dims, DiffField
    st = State(size=(16, 12, 10))
    og = st.og
    nu, kappa = 1e-6, 1e-7
    oc = K.Closure(nu=nu, kappa=kappa)
    U, Vp, W = average_dims(st.ou, (1, 2)), average_dims(st.ov, (1, 2)), average_dims(st.ow, (1, 2))
    up, vp, wp = DiffField(st.ou, U), DiffField(st.ov, Vp), DiffField(st.ow, W)
    ev = K.evaluate
    cells = [ev(K.shear_production_rate_ccc, og, "ccc", up, vp, wp, U, Vp, W),
             ev(K.viscous_dissipation_rate_ccc, og, "ccc", None, {"u": up, "v": vp, "w": wp, "b": st.ob}, {"closure": oc}),
             ev(K.turbulent_kinetic_energy_ccc, og, "ccc", st.ou, st.ov, st.ow, U, Vp, W)]
    jlo, jhi, klo, khi = window or (1, 12, 1, 10)
    Vol = 1.0 / (16 * 12 * 10)
    for ring in ([False, True] if window is None else [False]):
        got = class_sums(st, nu, kappa, jlo, jhi, klo, khi, ring=ring, profiles=(U, Vp, W))[:3]
        for name, c, g in zip(("SP", "EPS'", "TKE"), cells, got):
            win = c[:, jlo - 1:jhi, klo - 1:khi]
            want, scale = float(np.sum(win)) * Vol, float(np.sum(np.abs(win))) * Vol
            assert abs(g - want) <= 1e-12 * scale, (name, ring, g, want, scale)


def class_sums_stretched(st, nu, kappa, jlo, jhi, klo, khi, nu_e=None, kappa_scale=0.0):
    """The kernel's stretched-z GEN forms (`ocb_budget_sums_kernel<..., STR>`), optionally with an eddy-viscosity field nu_e inside
    the dissipation's items and kappa = kappa + kappa_scale x (two-point means of nu_e) on the tracer-gradient faces (`NUF`):
    per-plane weights dz_c (items of a cell plane), dz_f (fcf / cff strain edges), (G0 + G1) / 2 with G = class x dz_c (z-face items
    of K and w b), the by-parts advection split into the u / v equations' z parts (no metric left) and the w equation's parts with
    the dz_c-weighted transports."""
    og = st.og
    Nx, Ny, Nz = og.N
    H = og.H[0]
    rdx, rdy = Nx / 1.0, Ny / 1.0
    I = np.arange(1, Nx + 1)[:, None, None]
    J = np.arange(jlo, jhi + 2)[None, :, None]
    N = np.arange(klo, khi + 3)[None, None, :]
    at = lambda f, di=0, dj=0, dk=0: f.data[I + di + H - 1, J + dj + H - 1, N + dk + H - 1]
    lo_, hi_ = 1 - og.H[2], Nz + og.H[2]
    zc = lambda dk=0: np.asarray(og.spacing(2, "c", np.clip(N[0, 0, :] + dk, lo_, hi_)), dtype=np.float64)[None, None, :]
    zf = lambda dk=0: np.asarray(og.spacing(2, "f", np.clip(N[0, 0, :] + dk, lo_, hi_)), dtype=np.float64)[None, None, :]
    c0, c1, c2, dzf0, rzf0, rzf1, rzc1 = zc(0), zc(-1), zc(-2), zf(0), 1.0 / zf(0), 1.0 / zf(-1), 1.0 / zc(-1)
    u, v, w, b, p = st.ou, st.ov, st.ow, st.ob, st.op
    inw = lambda a, lo, hi: ((a >= lo) & (a <= hi)).astype(np.float64)
    al, be, an = inw(J, jlo, jhi), inw(J - 1, jlo, jhi), inw(J + 1, jlo, jhi)
    g0, g1, g2 = inw(N, klo, khi), inw(N - 1, klo, khi), inw(N - 2, klo, khi)
    G0, G1, G2 = g0 * c0, g1 * c1, g2 * c2
    hy, hyn, hz = 0.5 * (al + be), 0.5 * (al + an), 0.5 * (g0 + g1)
    hzK, rho0, rho1 = 0.5 * (G0 + G1), 0.5 * (G0 + G1) * rzf0, 0.5 * (G1 + G2) * rzf1
    uc, us, up_, ue = at(u), at(u, dj=-1), at(u, dk=-1), at(u, di=1)
    vc, vw, vp, vn, vs = at(v), at(v, di=-1), at(v, dk=-1), at(v, dj=1), at(v, dj=-1)
    wc, ww, ws, wp = at(w), at(w, di=-1), at(w, dj=-1), at(w, dk=-1)
    bc, bw, bs, bp, be_, bn = at(b), at(b, di=-1), at(b, dj=-1), at(b, dk=-1), at(b, di=1), at(b, dj=1)
    pc, pw, ps, pp = at(p), at(p, di=-1), at(p, dj=-1), at(p, dk=-1)
    duy, dvx, duz, dwx, dvz, dwy = uc - us, vc - vw, uc - up_, wc - ww, vc - vp, wc - ws
    ax, ay, dwz = ue - uc, vn - vc, wc - wp
    S12, S13, S23 = dvx * rdx + duy * rdy, dwx * rdx + duz * rzf0, dwy * rdy + dvz * rzf0
    P12, P13, P23 = (vw + vc) * (us + uc), (ww + wc) * (up_ + uc), (ws + wc) * (vp + vc)
    P13w, P23w = (ww + wc) * (c1 * up_ + c0 * uc), (ws + wc) * (c1 * vp + c0 * vc)
    Uu, Vv, Ww, VvS = (uc + ue) ** 2, (vc + vn) ** 2, (wp + wc) ** 2, (vs + vc) ** 2
    # viscosities / diffusivities of the items (one everywhere without a field: the constants are applied at the end)
    one = np.ones_like(uc)
    nffc = nfcf = ncff = ncc = nccp = one
    kx = ky = kz = kxe = kyn = one
    if nu_e is not None:
        ncn, nw_, ns_, nsw, npv, npw, nps = at(nu_e), at(nu_e, di=-1), at(nu_e, dj=-1), at(nu_e, di=-1, dj=-1), at(nu_e, dk=-1), at(nu_e, di=-1, dk=-1), at(nu_e, dj=-1, dk=-1)
        XN, YN, XNp, YNp = nw_ + ncn, ns_ + ncn, npw + npv, nps + npv
        nffc, nfcf, ncff = nu + 0.25 * ((nsw + ns_) + XN), nu + 0.25 * (XNp + XN), nu + 0.25 * (YNp + YN)
        ncc, nccp = nu + ncn, nu + npv
        hs = 0.5 * kappa_scale
        kx, ky, kz = kappa + hs * XN, kappa + hs * YN, kappa + hs * (npv + ncn)
    S = np.sum
    aEe = S(hy * G0 * nffc * S12 ** 2 + al * (hz * dzf0) * nfcf * S13 ** 2 + hy * (hz * dzf0) * ncff * S23 ** 2)
    aEx, aEy, aEz = S(al * G0 * ncc * ax ** 2), S(al * G0 * ncc * ay ** 2), S(al * g1 * rzc1 * nccp * dwz ** 2)
    aAe = S(G0 * P12 * (rdx * hy * dvx + rdy * (al * uc - be * us)) + al * P13w * (rdx * rho0) * dwx + al * P13 * (g0 * uc - g1 * up_) +
            P23w * (rdy * rho0) * (al * wc - be * ws) + hy * P23 * (g0 * vc - g1 * vp))
    aAx = S(al * G0 * Uu * ax)
    aAy = S(G0 * Vv * (hyn * vn - hy * vc)) + S((J == jlo) * G0 * VvS * 0.5 * vc)
    aAz = S(al * Ww * (rho0 * wc - rho1 * wp))
    aK = S(al * G0 * uc ** 2 + hy * G0 * vc ** 2 + al * hzK * wc ** 2)
    aPx, aPy, aPz = S(al * G0 * uc * (pc - pw)), S(hy * G0 * vc * (pc - ps)), S(al * rho0 * wc * (pc - pp))
    aW = S(al * hzK * wc * (bp + bc))
    dbx, dby, dbz = bc - bw, bc - bs, bc - bp
    aCx, aCy, aCz = S(al * G0 * kx * dbx ** 2), S(hy * G0 * ky * dby ** 2), S(al * (hz * rzf0) * kz * dbz ** 2)
    # dz_c x the Laplacian (sign of div q) of the cell below the face of step n
    bpw, bpe, bps, bpn, bpp = at(b, di=-1, dk=-1), at(b, di=1, dk=-1), at(b, dj=-1, dk=-1), at(b, dj=1, dk=-1), at(b, dk=-2)
    if nu_e is None:
        kxp = kxep = kyp = kynp = kzp = one
    else:
        hs = 0.5 * kappa_scale
        npe, npn, npp = at(nu_e, di=1, dk=-1), at(nu_e, dj=1, dk=-1), at(nu_e, dk=-2)
        kxp, kxep, kyp, kynp, kzp = kappa + hs * (npw + npv), kappa + hs * (npv + npe), kappa + hs * (nps + npv), kappa + hs * (npv + npn), kappa + hs * (npp + npv)
    lap = c1 * (rdx ** 2 * (kxp * (bp - bpw) - kxep * (bpe - bp)) + rdy ** 2 * (kyp * (bp - bps) - kynp * (bpn - bp))) + (kzp * (bp - bpp) * rzf1 - kz * dbz * rzf0)
    aD = S(al * g1 * bp * lap)
    V = 1.0 / (rdx * rdy)
    fn, fk = (1.0, 1.0) if nu_e is not None else (nu, kappa)
    return [V * 0.5 * aK, V * -0.25 * (aAe + rdx * aAx + rdy * aAy + aAz), V * (rdx * aPx + rdy * aPy + aPz), V * 0.5 * aW,
            V * fn * (aEe + 2.0 * (rdx ** 2 * aEx + rdy ** 2 * aEy + aEz)), V * 2.0 * fk * (rdx ** 2 * aCx + rdy ** 2 * aCy + aCz), V * 2.0 * fk * aD]


@pytest.mark.parametrize("window", [None, (3, 9, 1, 10), (1, 1, 2, 2), (12, 12, 10, 10), (1, 12, 4, 7), (2, 11, 1, 1)], ids=str)
@pytest.mark.parametrize("closure", ["scalar", "smagorinsky_like"])
def test_stretched_class_sums_reproduce_the_windowed_cell_integrals(window, closure):
    """The stretched-z forms of the class sums (and their eddy-viscosity / eddy-diffusivity variant) against the oracle's per-cell
    values times their cell volumes, on the reference's stretched test grid (test/test_utils.jl:17-24)."""
    st = State(size=(16, 12, 10), stretched=True)
    nu, kappa = 1e-6, 1e-7
    if closure == "scalar":
        oc, nu_e, nu0, ka0, ks = K.Closure(nu=nu, kappa=kappa), None, nu, kappa, 0.0
    else:
        oc, nu_e, nu0, ka0, ks = st.closures()[closure][0], st.onu, 0.0, 0.0, 1.0
    cells = _budget_cells(st, oc)
    ii, jj, kk = st.og.indices("ccc")
    vol = np.broadcast_to(st.og.V("ccc", ii, jj, kk), cells[0].shape)
    jlo, jhi, klo, khi = window or (1, 12, 1, 10)
    got = class_sums_stretched(st, nu0, ka0, jlo, jhi, klo, khi, nu_e=nu_e, kappa_scale=ks)
    for name, c, g in zip("K ADV PR WB EPS CHI CDF".split(), cells, got):
        win, vw = c[:, jlo - 1:jhi, klo - 1:khi], vol[:, jlo - 1:jhi, klo - 1:khi]
        want, scale = float(np.sum(win * vw)), float(np.sum(np.abs(win) * vw))
        assert abs(g - want) <= 1e-12 * scale, (name, g, want, scale)
