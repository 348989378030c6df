"""One-off wide parity fuzz on a GPU box: random small configurations, CUDA solve vs the CPU oracle.

    python tools/fuzz_parity.py [first_seed] [count] [max_seconds]

Prints one JSON line: solves checked, bit-identical solves, max |dVaR|, per-family mismatch list.
"""
import json
import sys
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent
for p in (REPO / "copula-msm-and-copula-garch-var_b200", REPO, REPO / "tests"):
    sys.path.insert(0, str(p))
from test_gpu_random import _random_case          # noqa: E402
from cvar_b200.backend import VarPlan              # noqa: E402
from oracle import var_oracle as vo                # noqa: E402

first = int(sys.argv[1]) if len(sys.argv) > 1 else 100
count = int(sys.argv[2]) if len(sys.argv) > 2 else 300
max_seconds = float(sys.argv[3]) if len(sys.argv) > 3 else float("inf")   # stop early (and still report) after this long
import time                                        # noqa: E402
t_start = time.perf_counter()
seeds_done = 0
solves = exact = 0
worst = 0.0
bad = []
for seed in range(first, first + count):
    if time.perf_counter() - t_start > max_seconds:
        break
    seeds_done += 1
    inp, alphas = _random_case(seed)
    with VarPlan(inp) as plan:
        res = plan.solve(inp.day_params(), alphas, ptf_mean=inp.ptf_mean)
    for k, a in enumerate(alphas):
        tr = vo.calc_var(inp, a)
        nan = np.isnan(res.var[k]) & np.isnan(tr.var)
        d = np.abs(np.where(nan, 0.0, res.var[k] - tr.var))
        if not np.array_equal(np.isnan(res.var[k]), np.isnan(tr.var)) or res.iterations[k] != tr.iterations:
            bad.append({"seed": seed, "alpha": a, "why": "nan pattern / iteration count"})
            continue
        solves += inp.T
        exact += int(np.sum((res.var[k] == tr.var) | nan))
        worst = max(worst, float(d.max()))
        if d.max() > 0:
            bad.append({"seed": seed, "alpha": a, "copula": inp.copula, "marginal": inp.marginal, "n": inp.n, "max_abs": float(d.max())})
print(json.dumps({"first_seed": first, "cases": seeds_done, "solves": solves, "bit_identical": exact, "max_abs_dvar": worst, "not_identical": bad}))
