"""The parity case table: every K1 diagnostic as (name, package constructor, oracle evaluation, location).
Used by the host-side evaluator check (CPU container) and by the GPU parity tests (through the C ABI)."""
from __future__ import annotations

import numpy as np

from helpers import ocb, K, State, OConst, OZero, DiffField, average_dims, mirror

KE, TKE, TV, FD = ocb.KineticEnergyEquation, ocb.TurbulentKineticEnergyEquation, ocb.TracerVarianceEquation, ocb.FlowDiagnostics


TILT = (float(np.sin(0.3)), 0.0, float(np.cos(0.3)))


def build_cases(st: State, closure_name="scalar", mean="profile"):
    """-> list of (name, package op, oracle values [i,j,k], location string)."""
    oc, pc = st.closures()[closure_name]
    f_cor = (1e-5, 2e-5, 1e-4)
    m = st.model(closure=pc, coriolis=ocb.ConstantCartesianCoriolis(*f_cor))
    og, of = st.og, st.ofields()
    par = {"closure": oc}
    ev = K.evaluate

    # mean flow: horizontal-mean profiles U(z), V(z), W(z) (reduced fields), numbers, or none
    if mean == "profile":
        oU, oV, oW = average_dims(st.ou, (1, 2)), average_dims(st.ov, (1, 2)), average_dims(st.ow, (1, 2))
        U, V, W = mirror(oU, st.pg), mirror(oV, st.pg), mirror(oW, st.pg)
    elif mean == "number":
        oU, oV, oW = OConst(0.3), OConst(-0.2), OConst(0.0)
        U, V, W = 0.3, -0.2, 0.0
    else:
        oU, oV, oW = OZero(), OZero(), OZero()
        U = V = W = None
    oup, ovp, owp = DiffField(st.ou, oU), DiffField(st.ov, oV), DiffField(st.ow, oW)
    ofp = {"u": oup, "v": ovp, "w": owp, "b": st.ob}
    sp_args = (oup, ovp, owp, oU, oV, oW)
    vdir = (0.0, 0.0, 1.0)
    tilt = TILT
    ro_bg = dict(dWdy_bg=1e-3, dVdz_bg=2e-3, dUdz_bg=-1e-3, dWdx_bg=5e-4, dUdy_bg=3e-3, dVdx_bg=-2e-3)
    oro = dict(ro_bg, fx=f_cor[0], fy=f_cor[1], fz=f_cor[2])
    ddir = (0.6, 0.0, 0.8)
    odepv = dict(dir_x=ddir[0], dir_y=ddir[1], dir_z=ddir[2], f_dir=sum(a * b for a, b in zip(f_cor, ddir)))
    m_tilt = st.model(closure=pc, buoyancy=ocb.BuoyancyTracer(gravity_unit_vector=tuple(-t for t in tilt)))

    cases = [
        ("KineticEnergy", KE.KineticEnergy(m), ev(K.kinetic_energy_ccc, og, "ccc", st.ou, st.ov, st.ow), "ccc"),
        ("KineticEnergyAdvection", KE.KineticEnergyAdvection(m), ev(K.ke_advection_ccc, og, "ccc", of), "ccc"),
        ("KineticEnergyPressureRedistribution", KE.KineticEnergyPressureRedistribution(m), ev(K.ke_pressure_ccc, og, "ccc", of, st.op), "ccc"),
        ("PotentialEnergyConversion", KE.PotentialEnergyConversion(m), ev(K.ke_buoyancy_ccc, og, "ccc", of, (0, 0, 1), st.ob), "ccc"),
        ("PotentialEnergyConversion_tilted", KE.PotentialEnergyConversion(m_tilt), ev(K.ke_buoyancy_ccc, og, "ccc", of, tilt, st.ob), "ccc"),
        ("DissipationRate", KE.DissipationRate(m), ev(K.viscous_dissipation_rate_ccc, og, "ccc", None, of, par), "ccc"),
        ("KineticEnergyStress", KE.KineticEnergyStress(m), ev(K.ke_stress_ccc, og, "ccc", oc, None, None, of, None), "ccc"),
        ("DissipationRate_perturbation", KE.DissipationRate(m, U=U, V=V, W=W) if U is not None else KE.DissipationRate(m),
         ev(K.viscous_dissipation_rate_ccc, og, "ccc", None, ofp, par), "ccc"),
        ("KineticEnergyIsotropicDissipationRate", KE.KineticEnergyIsotropicDissipationRate(m),
         ev(K.isotropic_viscous_dissipation_rate_ccc, og, "ccc", st.ou, st.ov, st.ow, par), "ccc"),
        ("TurbulentKineticEnergy", TKE.TurbulentKineticEnergy(m, U=U, V=V, W=W),
         ev(K.turbulent_kinetic_energy_ccc, og, "ccc", st.ou, st.ov, st.ow, oU, oV, oW), "ccc"),
        ("TKEIsotropicDissipationRate", TKE.TurbulentKineticEnergyIsotropicDissipationRate(m, U=U, V=V, W=W),
         ev(K.isotropic_viscous_dissipation_rate_ccc, og, "ccc", oup, ovp, owp, par), "ccc"),
        ("XShearProductionRate", TKE.XShearProductionRate(m, U=U, V=V, W=W), ev(K.shear_production_rate_x_ccc, og, "ccc", *sp_args), "ccc"),
        ("YShearProductionRate", TKE.YShearProductionRate(m, U=U, V=V, W=W), ev(K.shear_production_rate_y_ccc, og, "ccc", *sp_args), "ccc"),
        ("ZShearProductionRate", TKE.ZShearProductionRate(m, U=U, V=V, W=W), ev(K.shear_production_rate_z_ccc, og, "ccc", *sp_args), "ccc"),
        ("ShearProductionRate", TKE.ShearProductionRate(m, U=U, V=V, W=W), ev(K.shear_production_rate_ccc, og, "ccc", *sp_args), "ccc"),
        ("TracerVarianceDissipationRate", TV.TracerVarianceDissipationRate(m, "b"),
         ev(K.tracer_variance_dissipation_rate_ccc, og, "ccc", oc, None, None, st.ob), "ccc"),
        ("TracerVarianceDiffusion", TV.TracerVarianceDiffusion(m, "b"), ev(K.c_div_q_c, og, "ccc", oc, None, None, st.ob), "ccc"),
        ("RichardsonNumber", FD.RichardsonNumber(m), ev(K.richardson_number_ccf, og, "ccf", st.ou, st.ov, st.ow, st.ob, vdir), "ccf"),
        ("RichardsonNumber_tilted", FD.RichardsonNumber(m_tilt), ev(K.richardson_number_ccf, og, "ccf", st.ou, st.ov, st.ow, st.ob, tilt), "ccf"),
        ("RossbyNumber", FD.RossbyNumber(m, **ro_bg), ev(K.rossby_number_fff, og, "fff", st.ou, st.ov, st.ow, oro), "fff"),
        ("ErtelPotentialVorticity", FD.ErtelPotentialVorticity(m), ev(K.ertel_potential_vorticity_fff, og, "fff", st.ou, st.ov, st.ow, st.ob, *f_cor), "fff"),
        ("ErtelPotentialVorticity_thermal_wind", FD.ErtelPotentialVorticity(m, thermal_wind=True),
         ev(K.potential_vorticity_in_thermal_wind_fff, og, "fff", st.ou, st.ov, st.ob, f_cor[2]), "fff"),
        ("DirectionalErtelPotentialVorticity", FD.DirectionalErtelPotentialVorticity(m, ddir),
         ev(K.directional_ertel_potential_vorticity_fff, og, "fff", st.ou, st.ov, st.ow, st.ob, odepv), "fff"),
        ("StrainRateTensorModulus", FD.StrainRateTensorModulus(m), ev(K.strain_rate_tensor_modulus_ccc, og, "ccc", st.ou, st.ov, st.ow), "ccc"),
        ("VorticityTensorModulus", FD.VorticityTensorModulus(m), ev(K.vorticity_tensor_modulus_ccc, og, "ccc", st.ou, st.ov, st.ow), "ccc"),
        ("QVelocityGradientTensorInvariant", FD.QVelocityGradientTensorInvariant(m),
         ev(K.Q_velocity_gradient_tensor_invariant_ccc, og, "ccc", st.ou, st.ov, st.ow), "ccc"),
    ]
    # kinetic-energy forcing term (scope table 8f, row 2): space-dependent relaxation of each velocity, evaluated by the oracle's
    # ContinuousForcing at every stored point and handed to the package as forcing Fields (the functor path is a GPU test)
    relax = lambda x, y, z, t, q: -(0.5 + 0.25 * np.sin(2 * np.pi * x) * np.cos(2 * np.pi * y) + z) * q
    oforc, pforc = {}, {}
    for nm, loc, src in (("u", "fcc", st.ou), ("v", "cfc", st.ov), ("w", "ccf", st.ow)):
        F = K.ContinuousForcing(relax, loc, (nm,))
        oforc[nm] = F
        fld = type(src)(og, loc)
        H = og.H
        ii, jj, kk = (np.arange(1 - H[d], fld.data.shape[d] - H[d] + 1) for d in range(3))
        fld.data[...] = F(ii[:, None, None], jj[None, :, None], kk[None, None, :], og, None, of)
        pforc[nm] = mirror(fld, st.pg)
    m_forced = st.model(closure=pc)
    m_forced.forcing.u, m_forced.forcing.v, m_forced.forcing.w = pforc["u"], pforc["v"], pforc["w"]
    cases.append(("KineticEnergyForcing", KE.KineticEnergyForcing(m_forced), ev(K.ke_forcing_ccc, og, "ccc", oforc, None, of), "ccc"))
    # tracer-variance tendency (scope table 8f, row 2) of b with a space-dependent relaxation forcing handed over as a Field
    Fb = K.ContinuousForcing(relax, "ccc", ("b",))
    fb = type(st.ob)(og, "ccc")
    ii, jj, kk = (np.arange(1 - og.H[d], fb.data.shape[d] - og.H[d] + 1) for d in range(3))
    fb.data[...] = Fb(ii[:, None, None], jj[None, :, None], kk[None, None, :], og, None, of)
    m_forced.forcing.b = mirror(fb, st.pg)
    cases.append(("TracerVarianceTendency", TV.TracerVarianceTendency(m_forced, "b"),
                  ev(K.c_dt_c_ccc, og, "ccc", 1, "b", None, oc, None, None, None, None, {"u": st.ou, "v": st.ov, "w": st.ow}, {"b": st.ob},
                     None, None, None, Fb), "ccc"))
    # kinetic-energy tendency uᵢGᵢ (scope table 8f, row 2): Coriolis, buoyancy, viscous stress and the forcings above; once for a
    # model without a hydrostatic pressure anomaly (w feels buoyancy) and once with one (st.p stands in for pHY′)
    buoy = ((0.0, 0.0, 1.0), "b")
    vel = {"u": st.ou, "v": st.ov, "w": st.ow}
    for tag, opHY in (("", None), ("_pHY", st.op)):
        m_t = st.model(closure=pc, coriolis=ocb.ConstantCartesianCoriolis(*f_cor))
        m_t.forcing.u, m_t.forcing.v, m_t.forcing.w = pforc["u"], pforc["v"], pforc["w"]
        m_t.pressures.pHY = st.p if opHY is not None else None
        cases.append(("KineticEnergyTendency" + tag, KE.KineticEnergyTendency(m_t),
                      ev(K.ke_tendency_ccc, og, "ccc", None, f_cor, None, oc, None, None, None, buoy, None, vel, {"b": st.ob}, None, None, opHY, None,
                         oforc), "ccc"))
    # pass-through modules (scope table 8f, row 4): the terms of the momentum and tracer equations one by one, on the forced,
    # rotating, stratified model above with st.p standing in for the hydrostatic pressure anomaly
    m_p = st.model(closure=pc, coriolis=ocb.ConstantCartesianCoriolis(*f_cor), buoyancy=ocb.BuoyancyTracer(gravity_unit_vector=tuple(-t for t in tilt)))
    m_p.forcing.u, m_p.forcing.v, m_p.forcing.w, m_p.forcing.b = pforc["u"], pforc["v"], pforc["w"], m_forced.forcing.b
    m_p.pressures.pHY = st.p
    flds = {**vel, "b": st.ob}
    tbuoy = (tilt, "b")
    tend_args = (None, f_cor, None, oc, None, tbuoy, None, vel, {"b": st.ob}, None, None, st.op, None)
    mom = {"u": (ocb.UMomentumEquation, "fcc", K.div_Uu, K.x_dot_g_b_fcc, K.x_f_cross_U, K.ddx_fcc, K.div_tau_1, K.u_velocity_tendency),
           "v": (ocb.VMomentumEquation, "cfc", K.div_Uv, K.y_dot_g_b_cfc, K.y_f_cross_U, K.ddy_cfc, K.div_tau_2, K.v_velocity_tendency),
           "w": (ocb.WMomentumEquation, "ccf", K.div_Uw, K.z_dot_g_b_ccf, K.z_f_cross_U, None, K.div_tau_3, K.w_velocity_tendency)}
    for cn, (mod, loc, f_adv, f_buoy, f_cor_, f_pg, f_tau, f_tend) in mom.items():
        C = cn.upper()
        cases += [
            (f"{C}Advection", mod.Advection(m_p), ev(f_adv, og, loc, None, vel, vel[cn]), loc),
            (f"{C}BuoyancyAcceleration", mod.BuoyancyAcceleration(m_p), ev(f_buoy, og, loc, tilt, st.ob), loc),
            (f"{C}CoriolisAcceleration", mod.CoriolisAcceleration(m_p), ev(f_cor_, og, loc, f_cor, vel), loc),
            (f"{C}PressureGradient", mod.PressureGradient(m_p), ev(f_pg, og, loc, st.op) if f_pg else np.zeros(og.extent_of(loc)), loc),
            (f"{C}ViscousDissipation", mod.ViscousDissipation(m_p), ev(f_tau, og, loc, oc, None, None, flds, None), loc),
            (f"{C}Forcing", mod.Forcing(m_p), ev(oforc[cn], og, loc, None, flds), loc),
            (f"{C}Tendency", mod.Tendency(m_p), ev(f_tend, og, loc, *tend_args, oforc[cn]), loc),
        ]
    TE = ocb.TracerEquation
    Uo = (st.ou, st.ov, st.ow)
    cases += [
        ("TracerAdvection", TE.Advection(m_p, "b"), ev(K.div_Uc, og, "ccc", None, Uo, st.ob), "ccc"),
        ("TracerDiffusion", TE.Diffusion(m_p, "b"), ev(K.div_q_c, og, "ccc", oc, None, 1, st.ob, None, flds, None), "ccc"),
        ("TracerForcing", TE.Forcing(m_p, "b"), ev(Fb, og, "ccc", None, flds), "ccc"),
        ("TracerXDiffusiveFlux", TE.XDiffusiveFlux(m_p, "b"), ev(K.diffusive_flux_x, og, "fcc", oc, None, 1, st.ob), "fcc"),
        ("TracerYDiffusiveFlux", TE.YDiffusiveFlux(m_p, "b"), ev(K.diffusive_flux_y, og, "cfc", oc, None, 1, st.ob), "cfc"),
        ("TracerZDiffusiveFlux", TE.ZDiffusiveFlux(m_p, "b"), ev(K.diffusive_flux_z, og, "ccf", oc, None, 1, st.ob), "ccf"),
    ]
    # Centered(order = 4) advection (the scheme of every example of the reference, docs/examples/*.jl): the KE advection term, the
    # KE and tracer-variance tendencies, and the momentum / tracer advection terms
    m4 = st.model(closure=pc, coriolis=ocb.ConstantCartesianCoriolis(*f_cor), buoyancy=ocb.BuoyancyTracer(gravity_unit_vector=tuple(-t for t in tilt)))
    m4.advection = "Centered(order=4)"
    m4.forcing.u, m4.forcing.v, m4.forcing.w, m4.forcing.b = pforc["u"], pforc["v"], pforc["w"], m_forced.forcing.b
    o4 = "Centered(order=4)"
    cases += [
        ("KineticEnergyAdvection_o4", KE.KineticEnergyAdvection(m4), ev(K.ke_advection_ccc, og, "ccc", of, o4), "ccc"),
        ("KineticEnergyTendency_o4", KE.KineticEnergyTendency(m4),
         ev(K.ke_tendency_ccc, og, "ccc", o4, f_cor, None, oc, None, None, None, tbuoy, None, vel, {"b": st.ob}, None, None, None, None, oforc), "ccc"),
        ("TracerVarianceTendency_o4", TV.TracerVarianceTendency(m4, "b"),
         ev(K.c_dt_c_ccc, og, "ccc", 1, "b", o4, oc, None, None, None, None, {"u": st.ou, "v": st.ov, "w": st.ow}, {"b": st.ob}, None, None, None, Fb), "ccc"),
        ("UAdvection_o4", ocb.UMomentumEquation.Advection(m4), ev(K.div_Uu, og, "fcc", o4, vel, st.ou), "fcc"),
        ("VAdvection_o4", ocb.VMomentumEquation.Advection(m4), ev(K.div_Uv, og, "cfc", o4, vel, st.ov), "cfc"),
        ("WAdvection_o4", ocb.WMomentumEquation.Advection(m4), ev(K.div_Uw, og, "ccf", o4, vel, st.ow), "ccf"),
        ("TracerAdvection_o4", TE.Advection(m4, "b"), ev(K.div_Uc, og, "ccc", o4, Uo, st.ob), "ccc"),
    ]
    # Smagorinsky-Lilly eddy viscosity computed by the library (scope table 8f, row 3), without and with the stratification factor
    for tag, Cb in (("", None), ("_Cb", 1.0)):
        sm = ocb.SmagorinskyLilly(C=0.16, Cb=Cb, Pr=1.0)
        m_s = st.model(closure=sm)
        cases.append(("SmagorinskyViscosity" + tag, sm.nu_e.operand,
                      ev(K.smagorinsky_nu_ccc, og, "ccc", st.ou, st.ov, st.ow, st.ob, 0.16, Cb or 0.0, 1.0), "ccc"))
    # available-potential-energy family (scope table 8f): z✶ and Ψδ come from the oracle's sorted reference state and are
    # handed to both sides as fields, so these cases test the pointwise / flux kernels, not the sort (tests/test_bpe_gpu.py)
    from oracle import bpe as OB
    from oracle.grid import OField
    ref = OB.reference_height(st.ob, "HeavisideIntegral")

    def ofield(values):
        f = OField(og, "ccc")
        f.interior[...] = values
        f.fill_halo()
        return f
    ozs, opsi = ofield(ref["zstar"]), ofield(ref["psi_delta"])
    oups = ofield(ev(K.upsilon_ccc, og, "ccc", ozs))
    pzs, ppsi, pups = mirror(ozs, st.pg), mirror(opsi, st.pg), mirror(oups, st.pg)
    KFO, SO = ocb.KernelFunctionOperation, ocb.SweepOperands
    cases += [
        ("AvailablePotentialEnergy", KFO("APE", st.pg, SO(b=st.b, aux={5: ppsi, 6: pzs}), name="AvailablePotentialEnergy"),
         ev(K.local_ape_ccc, og, "ccc", opsi, st.ob, ozs), "ccc"),
        ("BuoyancyDisplacementPotential", KFO("UPSILON", st.pg, SO(aux={6: pzs}), name="BuoyancyDisplacementPotential"),
         ev(K.upsilon_ccc, og, "ccc", ozs), "ccc"),
        ("AvailablePotentialEnergyDissipationRate", KFO("APE_EPS", st.pg, SO(b=st.b, aux={4: pups}, closure=pc), name="AvailablePotentialEnergyDissipationRate"),
         ev(K.ape_dissipation_rate_ccc, og, "ccc", oups, oc, None, None, st.ob), "ccc"),
    ]
    S, O = FD.StrainRateTensor(m), FD.VorticityTensor(m)
    T, Tc = FD.StressTensor(m), FD.StressTensor(m, collocate_diagonals=True)
    comp_args = {(1, 1): (st.ou,), (2, 2): (st.ov,), (3, 3): (st.ow,), (1, 2): (st.ou, st.ov), (1, 3): (st.ou, st.ow), (2, 3): (st.ov, st.ow)}
    comp_loc = {(1, 1): "ccc", (2, 2): "ccc", (3, 3): "ccc", (1, 2): "ffc", (1, 3): "fcf", (2, 3): "cff"}
    sub = {1: "₁", 2: "₂", 3: "₃"}
    for (i, j), args in comp_args.items():
        key = sub[i] + sub[j]
        cases.append((f"S{i}{j}", getattr(S, "S" + key), ev(K.strain_rate_tensor_component(i, j), og, comp_loc[(i, j)], *args), comp_loc[(i, j)]))
        if i != j:
            cases.append((f"O{i}{j}", getattr(O, "Ω" + key), ev(K.vorticity_tensor_component(i, j), og, comp_loc[(i, j)], *args), comp_loc[(i, j)]))
            cases.append((f"T{i}{j}", getattr(T, "τ" + key), ev(K.stress_tensor_component(i, j), og, comp_loc[(i, j)], *args), comp_loc[(i, j)]))
        else:
            nat = {1: "fcc", 2: "cfc", 3: "ccf"}[i]
            cases.append((f"T{i}{j}", getattr(T, "τ" + key), ev(K.stress_tensor_component(i, j), og, nat, *args), nat))
            cases.append((f"T{i}{j}c", getattr(Tc, "τ" + key), ev(K.stress_tensor_component(i, j, True), og, "ccc", *args), "ccc"))
    return cases


def integral_scale(og, values, loc):
    i, j, k = og.indices(loc)
    return float(np.sum(np.abs(values) * og.V(loc, i, j, k)))


def richardson_tolerance(st, vdir, want, rtol=1e-12):
    """Per-point absolute tolerance for Ri = ∂b/∂ẑ / (∂uₕ/∂ẑ)²: the denominator is a difference of nearly equal
    square roots, so its relative rounding error is eps * (sum of |samples|) / |difference| -- the tolerance is the
    contract's rtol plus that conditioning term (the largest |Ri| are exactly the worst-conditioned points)."""
    og = st.og
    i, j, k = og.indices("ccf")
    a = (st.ou, st.ov, st.ow, vdir)
    uh = lambda ii, jj, kk: np.abs(K.uh_norm_ccc(ii, jj, kk, og, *a))
    rdx, rdy = 1.0 / og.Dx("ccc", i, j, k), 1.0 / og.Dy("ccc", i, j, k)
    rdz = 1.0 / og.Dz("ccf", i, j, k)
    mag = (abs(vdir[2]) * rdz * (uh(i, j, k) + uh(i, j, k - 1)) + abs(vdir[0]) * rdx * 0.5 * (uh(i + 1, j, k) + uh(i - 1, j, k))
           + abs(vdir[1]) * rdy * 0.5 * (uh(i, j + 1, k) + uh(i, j - 1, k)))
    dudx = K.ix_caa(i, j, k, og, K.ddx_fcc, K.uh_norm_ccc, *a)
    dudy = K.iy_aca(i, j, k, og, K.ddy_cfc, K.uh_norm_ccc, *a)
    dudz = K.ddz_ccf(i, j, k, og, K.uh_norm_ccc, *a)
    total = np.abs(dudx * vdir[0] + dudy * vdir[1] + dudz * vdir[2])
    # numerator: the projected buoyancy gradient is also a sum of differences that can cancel
    ab = lambda ii, jj, kk: np.abs(st.ob[ii, jj, kk])
    magb = (abs(vdir[2]) * rdz * (ab(i, j, k) + ab(i, j, k - 1))
            + abs(vdir[0]) * rdx * 0.25 * (ab(i + 1, j, k) + ab(i - 1, j, k) + ab(i + 1, j, k - 1) + ab(i - 1, j, k - 1))
            + abs(vdir[1]) * rdy * 0.25 * (ab(i, j + 1, k) + ab(i, j - 1, k) + ab(i, j + 1, k - 1) + ab(i, j - 1, k - 1)))
    dbdx = K.ixz_caf(i, j, k, og, K.ddx_fcc, st.ob)
    dbdy = K.iyz_acf(i, j, k, og, K.ddy_cfc, st.ob)
    dbdz = K.ddz_ccf(i, j, k, og, st.ob)
    totalb = np.abs(dbdx * vdir[0] + dbdy * vdir[1] + dbdz * vdir[2])
    with np.errstate(divide="ignore", invalid="ignore"):
        cond = np.where(total > 0, mag / total, np.inf) * 2 + np.where(totalb > 0, magb / totalb, np.inf)
    eps = np.finfo(og.dtype).eps
    with np.errstate(invalid="ignore"):
        tol = np.abs(want) * (rtol + 16 * eps * cond)
    # want == 0 with a vanishing denominator or numerator (the wall faces: both z differences are exactly zero there) gives
    # 0 * inf above; such points must come out as exactly the oracle's value
    return np.where(np.isnan(tol), 0.0, tol)


def check_case(st, name, got, want, rtol=1e-12):
    """Elementwise parity of one case: |got - want| <= rtol * max|want| (signed transport terms cancel locally, so
    the scale is the field's magnitude); Richardson numbers get their conditioning-aware per-point bound."""
    from helpers import assert_close
    finite = np.isfinite(want)
    assert finite.mean() > 0.9, name
    got = np.asarray(got, dtype=np.float64)
    # wherever the oracle is not finite (0/0 or x/0 of the reference formula) the library must be non-finite in the same way
    assert np.array_equal(np.isnan(want), np.isnan(got)) and np.array_equal(np.isposinf(want), np.isposinf(got)) and \
        np.array_equal(np.isneginf(want), np.isneginf(got)), f"{name}: non-finite points differ from the oracle's"
    if name.startswith("RichardsonNumber"):
        tol = richardson_tolerance(st, TILT if name.endswith("tilted") else (0.0, 0.0, 1.0), want, rtol)
        err = np.abs(np.where(finite, got, 0.0) - np.where(finite, want, 0.0))
        big = finite & np.isinf(tol)           # infinitely ill-conditioned (zero denominator difference with Ri != 0): unbounded
        bad = finite & ~big & (err > tol)      # every other point, the wall faces (tol = 0: exact) included
        assert big.mean() < 0.01, name
        assert not bad.any(), f"{name}: {int(bad.sum())} points beyond tolerance, worst err {float(np.max(err[bad])):.3g}"
        return
    g, w_ = np.where(finite, got, 0.0), np.where(finite, want, 0.0)
    assert_close(g, w_, rtol, name)
    if name.split(" ")[0] in ELEMENTWISE and rtol <= 1e-10:
        # sums of squares (no sign cancellation between terms): ELEMENTWISE relative bound, with a floor 1000 x below the
        # field-scale bound used for the signed transport terms
        scale = max(float(np.max(np.abs(w_))), np.finfo(np.float64).tiny)
        tol = rtol * np.abs(w_) + 1e-3 * rtol * scale
        err = np.abs(g - w_)
        bad = err > tol
        assert not bad.any(), f"{name}: {int(bad.sum())} points beyond the elementwise bound, worst ratio {float(np.max(err / tol)):.3g}"


# positive-definite diagnostics checked elementwise
ELEMENTWISE = {"KineticEnergy", "TurbulentKineticEnergy", "DissipationRate", "DissipationRate_perturbation",
               "KineticEnergyIsotropicDissipationRate", "TKEIsotropicDissipationRate", "TracerVarianceDissipationRate",
               "StrainRateTensorModulus", "VorticityTensorModulus", "T11", "T22", "T33", "T11c", "T22c", "T33c",
               "SmagorinskyViscosity"}
