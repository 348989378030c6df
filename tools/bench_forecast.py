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

# ---- returns -> VaR on the device (BASELINE configs[2] shape): one upload of the return series, MSM filter, solve,
# finalize, one download of the VaR vector; timed with CUDA events around the whole chain
import torch                                                       # noqa: E402
from cvar_b200.backend import VarPlan                              # noqa: E402
from cvar_b200.distributed import var_from_returns_sharded         # noqa: E402
from cvar_b200.inputs import make_inputs                           # noqa: E402

inp = make_inputs("student", "mixture", 2048, rho=0.6, nu=5.3, probs=pbs, sigma_states=sig)
pinned = torch.from_numpy(series).pin_memory()
host_var = torch.empty((1, T), dtype=torch.float64).pin_memory()
with VarPlan(inp) as plan:
    want = plan.solve(pbs, [0.01])
    producer = lambda r: fc.msm_forecast_device(r, prm, k, N)[0]   # noqa: E731
    times = []
    for it in range(8):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        var, case, iters = var_from_returns_sharded(plan, producer, pinned, N, [0.01])
        host_var.copy_(var, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    ms = float(np.median(times[3:]))
print(json.dumps({"pipeline": "returns -> MSM(k=8) state filter -> Student-copula VaR (c3 shape)", "days": T, "grid": "2048x2048",
                  "window_length": N, "ms_per_batch_e2e": ms, "days_per_s_e2e": T / (ms * 1e-3),
                  "h2d_bytes": int(pinned.numel() * 8), "d2h_bytes": int(host_var.numel() * 8),
                  "bit_identical_to_host_forecast_plus_host_solve": bool(host_var.numpy().tobytes() == want.var.tobytes())}))
