"""Golden vectors of the forecast producers (MSM rolling-window state filter, GARCH one-step forecast) from the
UNMODIFIED reference.  Run in the build container only:  python tests/golden/make_golden_forecast.py"""
import os
import pickle
import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REPO = HERE.parent.parent
sys.path.insert(0, str(REPO / "copula-msm-and-copula-garch-var_b200"))
from cvar_b200 import synthetic as syn                      # noqa: E402

REFERENCE = Path("/root/reference")


def build_cases():
    cases = []
    for name, k, (m0, sbar, b, gamma), N, T, seed in [
        ("msm_k2", 2, (0.4, 1.1, 3.0, 0.3), 60, 6, 21), ("msm_k3", 3, (0.55, 1.4, 5.0, 0.2), 80, 5, 22),
        ("msm_k4", 4, (0.35, 0.9, 2.5, 0.1), 120, 5, 23), ("msm_k5", 5, (0.6, 1.2, 2.0, 0.05), 90, 4, 24),
        ("msm_k8", 8, (0.4, 1.1, 3.0, 0.3), 150, 3, 25), ("msm_k8_m0_above_1", 8, (1.3, 1.0, 2.0, 0.15), 100, 3, 26),
    ]:
        series = syn.msm_simulate_returns(T + N - 1, k, m0, sbar, b, gamma, seed)
        cases.append(dict(kind="msm", name=name, k=k, m0=m0, sigma_bar=sbar, b=b, gamma=gamma, N=N, T=T, series=series))
    rng = np.random.default_rng(31)
    for name, omega, alpha, beta, N, T in [
        ("garch_11", 0.02, [0.09], [0.89], 200, 8), ("garch_21", 0.03, [0.05, 0.04], [0.88], 150, 6),
        ("garch_12", 0.05, [0.1], [0.5, 0.3], 150, 6), ("garch_33", 0.04, [0.05, 0.03, 0.02], [0.4, 0.3, 0.1], 120, 5),
    ]:
        series = rng.standard_normal(T + N - 1) * 1.1
        cases.append(dict(kind="garch", name=name, omega=omega, alpha=alpha, beta=beta, N=N, T=T, series=series))
    for name, a, l, q, N, T, seed in [("kalman_a", 0.97, 0.0, 0.15, 200, 6, 41), ("kalman_b", 0.9, 0.3, 0.3, 150, 5, 42),
                                       ("kalman_c", 0.99, -0.4, 0.05, 300, 4, 43)]:
        rng = np.random.default_rng(seed)
        x, logvol = l, []
        for _ in range(T + N - 1):
            x = a * (x - l) + l + q * rng.standard_normal()
            logvol.append(x)
        series = np.exp(np.array(logvol)) * rng.standard_normal(T + N - 1)
        cases.append(dict(kind="kalman", name=name, a=a, l=l, q=q, N=N, T=T, series=series))
    return cases


def main():
    if not REFERENCE.exists():
        raise SystemExit("the reference is not mounted at /root/reference")
    cases = build_cases()
    with tempfile.TemporaryDirectory() as tmp:
        tmp = Path(tmp)
        stubs = tmp / "stubs"
        (stubs / "matplotlib").mkdir(parents=True)
        (stubs / "yfinance.py").write_text("")
        (stubs / "matplotlib" / "__init__.py").write_text("")
        (stubs / "matplotlib" / "pyplot.py").write_text("")
        fin, fout = tmp / "in.pkl", tmp / "out.pkl"
        with open(fin, "wb") as f:
            pickle.dump(cases, f)
        env = dict(os.environ, PYTHONPATH=f"{stubs}:{REFERENCE}")
        subprocess.run([sys.executable, str(HERE / "_ref_forecast_worker.py"), str(fin), str(fout)], check=True, env=env, cwd=tmp)
        with open(fout, "rb") as f:
            results = pickle.load(f)
    blob = {}
    for c in cases:
        for key, val in c.items():
            if key not in ("name", "kind"):
                blob[f"{c['name']}__{key}"] = np.asarray(val)
        blob[f"{c['name']}__ref"] = results[c["name"]]
    np.savez_compressed(HERE / "forecast_producers.npz", **blob)
    print("wrote forecast_producers.npz with", [c["name"] for c in cases])


if __name__ == "__main__":
    main()
