import os, sys, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ocb_loader
ocb = ocb_loader.load()
from bench import make_state
def timed(fn, warm=2, reps=10):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
dev = torch.device("cuda", 0)
n = 512
grid, m = make_state(ocb, torch, (n, n, n), dev)
m.coriolis = ocb.FPlane(1e-4)
U, V, W = (ocb.average_xy(f) for f in (m.velocities.u, m.velocities.v, m.velocities.w))
FD, KE, TKE = ocb.FlowDiagnostics, ocb.KineticEnergyEquation, ocb.TurbulentKineticEnergyEquation
epv = FD.ErtelPotentialVorticity(m)
trio = [TKE.ShearProductionRate(m, U=U, V=V, W=W), KE.DissipationRate(m, U=U, V=V, W=W), KE.KineticEnergyPressureRedistribution(m)]
for name, ops in (("all four", [epv] + trio), ("EPV alone", [epv]), ("SP+eps'+PR", trio)):
    grp = ocb.FusedDiagnostics(*[ocb.Integral(o) for o in ops])
    ms = timed(grp.plan.execute)
    print(json.dumps({"name": name, "ms": round(ms, 4), "variant": grp.plan.variant, "launches": grp.plan.n_launches}))
