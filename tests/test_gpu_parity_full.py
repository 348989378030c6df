"""Parity at the FULL sizes of the BASELINE configurations, on a sample of days spread over the whole batch.

The GPU solves the complete batch (its iteration count K is a property of the whole batch, quirk Q7); the CPU oracle then
solves the sampled days with that K on the box's host cores.  Bar (north_star): |dVaR| <= 1e-7 in return units and equal
exceedance counts; the solved quantile is a dyadic midpoint, so the comparison is in fact bit for bit.

  c2  Student-t + GARCH,      1024^2, 1000 days x 2 alphas : 128 days
  c3  Student-t + MSM (q=9),  2048^2, 1000 days            : 128 days
  c4  Plackett + Kalman,      2048^2, 1000 days x 2 alphas : 128 days
  c5  Student-t + MSM member of the 4096^2 sweep, 64 days x 2 alphas : 8 days
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CASES = [("c2", None, None, 128), ("c3", None, None, 128), ("c4", None, None, 128), ("c5_student_mixture", 64, 4096, 8)]


@pytest.mark.parametrize("name,T,n,sample_days", CASES, ids=[c[0] for c in CASES])
def test_full_size_sample_matches_the_oracle(cuda_device, name, T, n, sample_days):
    from bench import cpu_port_solve, make_pool          # the oracle runs in worker processes (one per host core)
    from cvar_b200 import synthetic as syn
    from cvar_b200.backend import VarPlan
    from cvar_b200.backtest import exceedances

    inp, alphas = syn.baseline_config(name, T=T, n=n)
    with VarPlan(inp, device=0) as plan:
        res = plan.solve(inp.day_params(), alphas, ptf_mean=inp.ptf_mean)
    assert not np.any(res.status), res.status
    sample = list(range(0, inp.T, max(1, inp.T // sample_days)))[:sample_days]
    workers = min(len(os.sched_getaffinity(0)), 32)
    pool = make_pool(workers)
    try:
        _, cpu = cpu_port_solve(inp, alphas, sample, [int(k) for k in res.iterations], pool, workers)
    finally:
        pool.shutdown()
    rng = np.random.default_rng(3)
    r_ptf = rng.standard_normal(len(sample)) * 1.3
    for k, a in enumerate(alphas):
        gpu = res.var[k][sample]
        assert np.max(np.abs(gpu - cpu[a])) <= 1e-7, (name, a)
        assert gpu.tobytes() == cpu[a].tobytes(), (name, a, float(np.max(np.abs(gpu - cpu[a]))))
        assert exceedances(gpu, r_ptf) == exceedances(cpu[a], r_ptf)
