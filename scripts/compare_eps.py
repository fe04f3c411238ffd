"""Dissipation rate alone at 512^3 Float64 on the three kernels that can evaluate it (tuning aid)."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ocb_loader
from bench import make_state
from scripts.bench_kernels import timed

ocb = ocb_loader.load()


def main():
    dev = torch.device("cuda", 0)
    n = 512
    grid, m = make_state(ocb, torch, (n, n, n), dev)
    KE, FD = ocb.KineticEnergyEquation, ocb.FlowDiagnostics
    eps, smod = KE.DissipationRate(m), FD.StrainRateTensorModulus(m)
    for name, members, env in (("eps integral, budget kernel", [ocb.Integral(eps)], {}),
                               ("eps field, budget kernel", [ocb.Field(eps)], {}),
                               ("eps + |S| integrals, velocity-gradient kernel", [ocb.Integral(eps), ocb.Integral(smod)], {}),
                               ("eps + |S| fields, velocity-gradient kernel", [ocb.Field(eps), ocb.Field(smod)], {}),
                               ("eps integral, generic kernel", [ocb.Integral(eps)], {"OCB_DISABLE_FUSED": "1"})):
        os.environ.update(env)
        try:
            grp = ocb.FusedDiagnostics(*members)
        finally:
            for k in env:
                os.environ.pop(k, None)
        ms = timed(grp.plan.execute)
        bpc = grp.plan.algorithmic_bytes / n ** 3
        print(f"{name:50s} variant {grp.plan.variant}  {ms:7.3f} ms  {n ** 3 / ms / 1e6:7.1f} Gcells/s  {bpc:4.0f} B/cell  "
              f"{bpc * n ** 3 / ms / 1e6 / 6542.4:5.3f} of HBM peak")


if __name__ == "__main__":
    main()
