"""Throughput of the GPU forecast producers at BASELINE scale next to the CPU port (one JSON line each).

    python tools/bench_forecast.py            # MSM k=8, T=1000 windows of N=1135 returns, 2 assets; GARCH(1,1) same shape
"""
import json
import sys
import time
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO / "copula-msm-and-copula-garch-var_b200"))
sys.path.insert(0, str(REPO))
from cvar_b200 import forecast as fc, synthetic as syn           # noqa: E402
from oracle import forecast_oracle as fo                          # noqa: E402  (CPU baseline only)

k, T, N = 8, 1000, 1135
prm = [fc.MsmParams(m0, sb, b, g) for (m0, sb, b, g, _) in syn.MSM_ASSETS]
series = np.array([syn.msm_simulate_returns(T + N - 1, k, p.m0, p.sigma_bar, p.b, p.gamma, 5 + i) for i, p in enumerate(prm)])
fc.msm_forecast(series, prm, k, N)                                # warm-up (context, module load)
best = min(fc.msm_forecast(series, prm, k, N)[2]["kernel_ms"] for _ in range(5))
t0 = time.perf_counter(); pbs, sig, info = fc.msm_forecast(series, prm, k, N); wall = time.perf_counter() - t0
sample = [0, T // 2, T - 1]
t0 = time.perf_counter()
cpu = np.array([fo.msm_forecast(series[0][w:w + N], k, prm[0].m0, prm[0].sigma_bar, prm[0].b, prm[0].gamma, N, 1)[0] for w in sample])
cpu_s = (time.perf_counter() - t0) / len(sample)
from cvar_b200.msm_layout import merge_states, msm_vol_states
want, _ = merge_states(msm_vol_states(k, prm[0].m0, prm[0].sigma_bar)[None, :], cpu[None, :, :])
err = float(np.max(np.abs(pbs[sample, 0, :] - want[:, 0, :])))
windows = 2 * T
print(json.dumps({"producer": "msm_state_filter", "k": k, "states": 2 ** k, "windows": windows, "window_length": N,
                  "kernel_ms": best, "windows_per_s_kernel": windows / (best * 1e-3), "host_call_s": wall,
                  "dense_flops_reference": windows * N * 2.0 * 4 ** k, "kronecker_flops": windows * N * (4.0 * k + 4) * 2 ** k,
                  "cpu_port_s_per_window": cpu_s, "speedup_vs_one_core": cpu_s * windows / (best * 1e-3),
                  "max_abs_err_vs_oracle_on_sample": err}))
rng = np.random.default_rng(3)
gser = rng.standard_normal((2, T + N - 1)) * 1.1
fc.garch_forecast(gser, [0.02, 0.03], [[0.09], [0.08]], [[0.89], [0.90]], N)
best = min(fc.garch_forecast(gser, [0.02, 0.03], [[0.09], [0.08]], [[0.89], [0.90]], N)[1]["kernel_ms"] for _ in range(5))
t0 = time.perf_counter(); [fo.garch_forecast_one(0.02, [0.09], [0.89], gser[0][w:w + N]) for w in range(20)]; cpu_s = (time.perf_counter() - t0) / 20
print(json.dumps({"producer": "garch_forecast", "windows": windows, "window_length": N, "kernel_ms": best,
                  "windows_per_s_kernel": windows / (best * 1e-3), "cpu_port_s_per_window": cpu_s,
                  "speedup_vs_one_core": cpu_s * windows / (best * 1e-3)}))
