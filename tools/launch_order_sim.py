"""Developer tool (CPU only): replay a launch of the solve kernel on a model of the SMs and compare launch orders.

Per-day costs come from the CPU oracle's cell counts of the c3 workload at a reduced grid (the cost of a solve is
~affine in the cells its strips cover: DESIGN.md section 6), the order keys are the portfolio-variance proxy of
order_key_kernel, the arranged ranks come from the shipped library (cvar_launch_order).  Model: every SM holds `per`
CTAs; a CTA advances at rate 1 when its SM is full and up to `alone` times faster when it is the only resident; a freed
slot takes the next CTA of the launch (block-index order).  Prints the time of the launch relative to the time the same
days take inside a long batch.

    python tools/launch_order_sim.py [--days 1000] [--per 2] [--alone 1.7]
"""
import argparse
import ctypes as C
import sys
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(REPO / "copula-msm-and-copula-garch-var_b200"), str(REPO)]

from cvar_b200 import _lib, synthetic as syn      # noqa: E402
from oracle import var_oracle as vo               # noqa: E402  (cost model only)


def replay(order, cost, sms, per, alone):
    n, nxt, t = len(order), 0, 0.0
    resident = [[] for _ in range(sms)]
    for _ in range(per):
        for m in range(sms):
            if nxt < n:
                resident[m].append(cost[order[nxt]])
                nxt += 1

    def rate(k):
        return 1.0 if k == per or per == 1 else 1.0 + (alone - 1.0) * (per - k) / (per - 1)

    while True:
        step = min((min(r) / rate(len(r)) for r in resident if r), default=None)
        if step is None:
            return t
        t += step
        freed = []
        for m, r in enumerate(resident):
            if not r:
                continue
            left = [w - step * rate(len(r)) for w in r]
            resident[m] = [w for w in left if w > 1e-9]
            freed += [m] * (len(left) - len(resident[m]))
        for m in freed:
            if nxt < n:
                resident[m].append(cost[order[nxt]])
                nxt += 1


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--days", type=int, default=1000)
    ap.add_argument("--sms", type=int, default=148)
    ap.add_argument("--per", type=int, default=2, help="resident CTAs per SM (2 at n = 2048, 4 at n = 1024)")
    ap.add_argument("--alone", type=float, default=1.7, help="speed of a CTA that has its SM to itself")
    args = ap.parse_args()

    inp, alphas = syn.baseline_config("c3", T=1000, n=256)
    cells = vo.calc_var(inp, alphas[0]).cells.astype(float)
    cost = cells / cells.mean() * 630e3 + 266e3          # cycles: cell loop + the per-day fixed part (phase profile, c3)
    v = np.einsum("tas,as->ta", inp.probs, inp.sigma_states ** 2)
    w = inp.weights
    key = w[0] ** 2 * v[:, 0] + w[1] ** 2 * v[:, 1] + 2 * inp.rho * w[0] * w[1] * np.sqrt(v[:, 0] * v[:, 1])
    if args.days != 1000:
        pick = np.random.default_rng(1).integers(0, 1000, args.days)
        cost, key = cost[pick], key[pick]

    slots = args.sms * args.per
    by_key = np.argsort(key, kind="stable")              # ascending variance: most expensive first
    ranks = np.empty(args.days, dtype=np.int32)
    rc = _lib.load().cvar_launch_order(args.days, slots, ranks.ctypes.data_as(C.POINTER(C.c_int32)))
    assert rc == 0
    long_cost = np.tile(cost, 8)
    per_batch = replay(np.argsort(-long_cost), long_cost, args.sms, args.per, args.alone) / 8
    for name, order in (("natural", np.arange(args.days)), ("most expensive first (proxy)", by_key),
                        ("most expensive first (true cost)", np.argsort(-cost)), ("arranged (proxy)", by_key[ranks])):
        print(f"{name:34s} {replay(order, cost, args.sms, args.per, args.alone) / per_batch:.3f} x the time inside a long batch")


if __name__ == "__main__":
    main()
