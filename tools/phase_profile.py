"""Per-phase cycle counts of the solve kernel from a -DCVAR_PROFILE_PHASES build (thread 0 of every CTA):
    CVAR_B200_LIB=build_exp/lib_prof.so python tools/phase_profile.py c3
"""
import sys
from pathlib import Path
import numpy as np
import torch
REPO = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(REPO / "copula-msm-and-copula-garch-var_b200"), str(REPO)]
from cvar_b200 import synthetic as syn
from cvar_b200.backend import VarPlan

name = sys.argv[1] if len(sys.argv) > 1 else "c3"
T = int(sys.argv[2]) if len(sys.argv) > 2 else None
inp, alphas = syn.baseline_config(name, T=T)
alphas = alphas[:1]
plan = VarPlan(inp, device=0)
d = torch.from_numpy(inp.day_params()).cuda()
T = inp.T
traj = torch.zeros((1, T, 2), dtype=torch.int32, device="cuda")
mass = torch.zeros((1, T), dtype=torch.float64, device="cuda")
cells = torch.zeros((1, T), dtype=torch.int64, device="cuda")
for _ in range(3):
    plan.solve_device(d, alphas, traj=traj, mass=mass, cells=cells)
torch.cuda.synchronize()
t = traj.cpu().numpy().astype(np.int64) & 0xffffffff
stage0 = t[0, :, 0] * 16
thin = t[0, :, 1] * 16
probe = mass.cpu().numpy()[0]
total = cells.cpu().numpy()[0]
dense = total - stage0 - thin - probe
print(f"{name}: per-day CTA cycles (mean over {T} days): total {total.mean():.0f}  stage0 {stage0.mean():.0f} ({100*stage0.mean()/total.mean():.1f} %)  "
      f"probes+brackets {probe.mean():.0f} ({100*probe.mean()/total.mean():.1f} %)  bisection k<10 {dense.mean():.0f} ({100*dense.mean()/total.mean():.1f} %)  "
      f"bisection k>=10 {thin.mean():.0f} ({100*thin.mean()/total.mean():.1f} %)")
