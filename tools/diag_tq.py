"""GPU diagnostic: device Student-t quantile (table and iterative) against SciPy, by region of u."""
import sys
from pathlib import Path
import numpy as np
from scipy import stats
REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO / "copula-msm-and-copula-garch-var_b200"))
from cvar_b200.backend import VarPlan
from cvar_b200.inputs import make_inputs

rng = np.random.default_rng(0)
regions = {
    "1e-17..1e-10": 10.0 ** rng.uniform(-17, -10, 3000),
    "1e-10..1e-4": 10.0 ** rng.uniform(-10, -4, 3000),
    "1e-4..0.05": 10.0 ** rng.uniform(-4, -1.3, 3000),
    "0.05..0.2": rng.uniform(0.05, 0.2, 3000),
    "0.2..0.45": rng.uniform(0.2, 0.45, 3000),
    "0.45..0.5": rng.uniform(0.45, 0.5, 3000),
    "0.5..1": rng.uniform(0.5, 1, 3000),
}
for nu in (2.01, 2.5, 4.0, 5.3, 30.0, 50.0):
    inp = make_inputs("student", "single", 64, nu=nu, sigma=np.ones((1, 2)))
    with VarPlan(inp) as plan:
        print(f"nu={nu}: table-vs-iterative max rel err {plan.info().tq_table_max_rel_err:.3e}")
        for name, u in regions.items():
            ref = stats.t.ppf(u, df=nu)
            fast, slow = plan.special(0, u), plan.special(1, u)
            sc = np.maximum(np.abs(ref), 0.1)
            ef, es = np.abs(fast - ref) / sc, np.abs(slow - ref) / sc
            print(f"   {name:14s} fast {ef.max():.2e} (u={u[ef.argmax()]:.3e})  slow {es.max():.2e} (u={u[es.argmax()]:.3e})"
                  f"  fast-vs-slow {np.max(np.abs(fast-slow)/sc):.2e}")
