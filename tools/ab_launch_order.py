"""Developer tool (GPU): time the solve launch of a workload under several launch-order settings in ONE process.

    python tools/ab_launch_order.py c3 c4 c2 [--steps 20] [--days 1000]

Settings are environment knobs read at plan creation (include/cvar.h): a plan per setting, same device inputs, L2 flushed
before every step, CUDA events around the solve launch (order kernels + solve kernel).  Also checks that every setting
gives bit-identical trajectories.
"""
import argparse
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch

REPO = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(REPO / "copula-msm-and-copula-garch-var_b200"), str(REPO)]

from cvar_b200 import synthetic as syn          # noqa: E402
from cvar_b200.backend import VarPlan           # noqa: E402

SETTINGS = {
    "natural": {"CVAR_ORDER_MIN_WAVES": "1000000"},
    "sorted": {},
    "sorted_from_1_wave": {"CVAR_ORDER_MIN_WAVES": "1"},
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("workloads", nargs="+")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--days", type=int, default=None)
    ap.add_argument("--settings", default=",".join(SETTINGS))
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    scratch = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for name in args.workloads:
        inp, alphas = syn.baseline_config(name, **({"T": args.days} if args.days else {}))
        d_day = torch.from_numpy(inp.day_params()).to(dev)
        ref = None
        for tag in args.settings.split(","):
            for k in ("CVAR_LAUNCH_ORDER", "CVAR_ORDER_MIN_WAVES"):
                os.environ.pop(k, None)
            os.environ.update(SETTINGS[tag])
            with VarPlan(inp, device=0) as plan:
                plan.reserve(inp.T)
                traj = torch.empty((len(alphas), inp.T, 2), dtype=torch.int32, device=dev)
                for _ in range(3):
                    plan.solve_device(d_day, alphas, traj=traj)
                ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
                for a, b in ev:
                    scratch.zero_()
                    a.record()
                    plan.solve_device(d_day, alphas, traj=traj)
                    b.record()
                torch.cuda.synchronize()
                ms = [a.elapsed_time(b) for a, b in ev]
                got = traj.cpu().numpy()
                same = True if ref is None else bool(np.array_equal(ref, got))
                ref = got if ref is None else ref
                print(json.dumps({"workload": name, "days": inp.T, "setting": tag, "ms_mean": float(np.mean(ms)),
                                  "ms_min": float(np.min(ms)), "ms_max": float(np.max(ms)),
                                  "solves_per_s": inp.T * len(alphas) / (np.mean(ms) * 1e-3), "same_traj_as_first": same}),
                      flush=True)


if __name__ == "__main__":
    main()
