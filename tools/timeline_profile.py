"""Developer tool (GPU): where and when every CTA of one solve launch ran, from a -DCVAR_PROFILE_TIMELINE build.

    tools/build_variant.sh $PWD/build_exp/lib_timeline.so -DCVAR_PROFILE_TIMELINE
    CVAR_B200_LIB=build_exp/lib_timeline.so python tools/timeline_profile.py c3 out_dir [settings]

Writes <out_dir>/timeline_<workload>_<setting>.npz with, per day: sm, launch position (block index), start and end
(globaltimer, ns, relative to the first start).  Prints the launch's span, the busy time of the SM slots, and how many
days each slot ran.  The build overwrites the results of the solve: never use it for anything else.
"""
import os
import sys
from pathlib import Path

import numpy as np
import torch

REPO = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(REPO / "copula-msm-and-copula-garch-var_b200"), str(REPO)]
from cvar_b200 import synthetic as syn          # noqa: E402
from cvar_b200.backend import VarPlan           # noqa: E402

SETTINGS = {
    "natural": {"CVAR_ORDER_MIN_WAVES": "1000000"},
    "sorted": {},
}

name = sys.argv[1] if len(sys.argv) > 1 else "c3"
out_dir = Path(sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/order")
settings = (sys.argv[3] if len(sys.argv) > 3 else "sorted,natural").split(",")
out_dir.mkdir(parents=True, exist_ok=True)
inp, alphas = syn.baseline_config(name)
alphas = alphas[:1]
d = torch.from_numpy(inp.day_params()).cuda()
T = inp.T
for tag in settings:
    for k in ("CVAR_LAUNCH_ORDER", "CVAR_ORDER_MIN_WAVES"):
        os.environ.pop(k, None)
    os.environ.update(SETTINGS[tag])
    with VarPlan(inp, device=0) as plan:
        traj = torch.zeros((1, T, 2), dtype=torch.int32, device="cuda")
        mass = torch.zeros((1, T), dtype=torch.float64, device="cuda")
        cells = torch.zeros((1, T), dtype=torch.int64, device="cuda")
        for _ in range(3):
            plan.solve_device(d, alphas, traj=traj, mass=mass, cells=cells)
        torch.cuda.synchronize()
        t = traj.cpu().numpy()[0].astype(np.int64) & 0xffffffff
        sm, pos = t[:, 0], t[:, 1]
        start = mass.cpu().numpy()[0]
        end = cells.cpu().numpy()[0].astype(np.float64)
        t0 = start.min()
        start, end = start - t0, end - t0
        np.savez(out_dir / f"timeline_{name}_{tag}.npz", sm=sm, pos=pos, start=start, end=end)
        span = end.max()
        dur = end - start
        per_sm = np.bincount(sm, minlength=148)
        first_wave = np.sort(start)[: min(T, 296)].max()
        print(f"{name} {tag}: span {span / 1e3:.1f} us, sum of CTA durations {dur.sum() / 1e3:.0f} us "
              f"(= {dur.sum() / span:.1f} slots busy on average), CTA duration min/mean/max {dur.min() / 1e3:.0f}/{dur.mean() / 1e3:.0f}/{dur.max() / 1e3:.0f} us, "
              f"days per SM min/max {per_sm.min()}/{per_sm.max()}, last start of the first 296: {first_wave / 1e3:.1f} us, "
              f"last start overall {start.max() / 1e3:.1f} us", flush=True)
