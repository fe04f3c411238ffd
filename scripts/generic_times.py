import os, sys, json, torch
sys.path.insert(0, '/root/repo')
os.environ["OCB_DISABLE_FUSED"] = "1"
import ocb_loader; ocb = ocb_loader.load()
from bench import make_state
from scripts.bench_kernels import timed
dev = torch.device("cuda", 0)
n = 512
grid, m = make_state(ocb, torch, (n, n, n), dev)
m.coriolis = ocb.FPlane(1e-4)
U, V, W = (ocb.average_xy(f) for f in (m.velocities.u, m.velocities.v, m.velocities.w))
FD, KE, TKE, TV = ocb.FlowDiagnostics, ocb.KineticEnergyEquation, ocb.TurbulentKineticEnergyEquation, ocb.TracerVarianceEquation
ops = {"EPV": FD.ErtelPotentialVorticity(m), "SP": TKE.ShearProductionRate(m, U=U, V=V, W=W), "EPS_pert": KE.DissipationRate(m, U=U, V=V, W=W),
       "PR": KE.KineticEnergyPressureRedistribution(m), "EPS": KE.DissipationRate(m), "ADV": KE.KineticEnergyAdvection(m), "KE": KE.KineticEnergy(m),
       "RI": FD.RichardsonNumber(m), "SMOD": FD.StrainRateTensorModulus(m), "CHI": TV.TracerVarianceDissipationRate(m, "b"), "Q": FD.QVelocityGradientTensorInvariant(m)}
for name, op in ops.items():
    grp = ocb.FusedDiagnostics(ocb.Integral(op))
    ms = timed(grp.plan.execute, warm=1, reps=3)
    print(name, round(ms, 3), "ms", round(n**3 / ms / 1e6, 2), "Gcells/s", flush=True)
