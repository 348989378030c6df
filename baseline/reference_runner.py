"""Running the UNMODIFIED reference (test and bench helper, never imported by the product).

`baseline/_ref` is what `baseline/install_reference.sh` pip-installs from the read-only checkout (git-ignored, travels to
the GPU box); in the build container the checkout itself, /root/reference, also works.  The reference's `utils` package
and this repository's mirror share their module names, so the reference always runs in a subprocess whose PYTHONPATH
puts it first.
"""
from __future__ import annotations

import os
import sys
from pathlib import Path

import json
import subprocess
import textwrap

REPO = Path(__file__).resolve().parent.parent
PKG_ROOT = REPO / "copula-msm-and-copula-garch-var_b200"


def reference_root() -> Path | None:
    for cand in (os.environ.get("CVAR_REFERENCE_PATH"), REPO / "baseline" / "_ref", "/root/reference"):
        if cand and (Path(cand) / "utils" / "calc_var_class.py").exists():
            return Path(cand)
    return None


def reference_env(stub_dir: Path | None = None, with_backend: bool = True) -> dict:
    """Environment of a subprocess that imports the reference's `utils` (first on the path), the stubs for the two
    modules the image lacks (yfinance, matplotlib) and, after them, this repository's `cvar_b200` package."""
    root = reference_root()
    if root is None:
        raise FileNotFoundError("no reference install: run baseline/install_reference.sh")
    stubs = root / "_stubs"
    if not stubs.exists():
        if stub_dir is None:
            raise FileNotFoundError("stub directory needed for a bare checkout")
        stubs = Path(stub_dir)
        (stubs / "matplotlib").mkdir(parents=True, exist_ok=True)
        (stubs / "yfinance.py").write_text("")
        (stubs / "matplotlib" / "__init__.py").write_text("")
        (stubs / "matplotlib" / "pyplot.py").write_text("")
    parts = [str(stubs), str(root)] + ([str(PKG_ROOT)] if with_backend else [])
    env = dict(os.environ, PYTHONPATH=os.pathsep.join(parts))
    env.pop("CVAR_BACKEND", None)
    return env


BUILD_OBJECT = '''
import numpy as np
from utils.factory import ValueAtRiskCalculationFactory          # the reference's
from utils.calc_var_class import ValueAtRiskCalcualtion          # the reference's


def build_reference_object(copula_type, estimation, n, copula_params, weights=(0.5, 0.5), ptf_mean=0.0, sigma=None,
                           probs=None, sigma_states=None, backend=None):
    """A real reference driver object with injected hot-path attributes (SURVEY App. C): no download, no fit."""
    m = (ValueAtRiskCalculationFactory.create_var_calculator(copula_type, estimation) if backend is None else
         ValueAtRiskCalculationFactory.create_var_calculator(copula_type, estimation, backend=backend))
    v = object.__new__(ValueAtRiskCalcualtion)
    v.VaRCalculationMethod = m
    v.num_points, v.weights, v.dim, v.ptf_mean = n, np.asarray(weights, float), 2, float(ptf_mean)
    if estimation == "msm":
        fbs, uvs = np.asarray(probs, float), np.asarray(sigma_states, float)
        v.out_sample_N = fbs.shape[0]
        dens, x, dx = m.compute_normal_densities(uvs, n)
        v.grids_generations_params = (dens, x, dx, m.create_vol_combinations(uvs))
        v.integrations_params_t = (fbs, m.compute_forecast_combinations(fbs))
        v.integrations_params_static = uvs
    else:
        sigma = np.asarray(sigma, float)
        v.out_sample_N = sigma.shape[0]
        dens, x, dx = m.compute_normal_densities(2, n)
        v.grids_generations_params = (dens, x, dx, np.zeros((1, 2)))
        v.integrations_params_t = [sigma]
        v.integrations_params_static = None
    v.copula_params = copula_params
    v.copula_function = m.copula_density
    v.unpack_copula_params = m.unpack_copula_params
    v.integrated_function = m.integrated_function
    return v
'''


TIME_CALC_VAR = BUILD_OBJECT + textwrap.dedent("""
    import contextlib, io, json, pickle, sys, time
    case = pickle.load(open(sys.argv[1], "rb"))
    v = build_reference_object(**case["object"])
    times, var = [], None
    for _ in range(case["repeats"]):
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()):          # the reference prints timings
            var = np.asarray(v.calc_var(obj_var=case["alpha"]))
        times.append(time.perf_counter() - t0)
    pickle.dump({"seconds": times, "var": var}, open(sys.argv[2], "wb"))
""")


def time_reference_calc_var(inp, alpha: float, repeats: int = 2, timeout: float = 1800.0, workdir=None) -> dict:
    """Run the unmodified reference's `calc_var(obj_var=alpha)` on the inputs of `inp` (cvar_b200.inputs.HotPathInputs)
    `repeats` times in one subprocess (first call = numba compilation + joblib worker start, later calls warm).
    Returns {"seconds": [...], "var": ndarray}.  Raises FileNotFoundError without a reference install."""
    import pickle
    import tempfile

    estimation = "msm" if inp.marginal == "mixture" else "garch"
    obj = dict(copula_type=inp.copula, estimation=estimation, n=int(inp.n), copula_params=inp.copula_params(),
               weights=tuple(float(w) for w in inp.weights), ptf_mean=float(inp.ptf_mean), sigma=inp.sigma, probs=inp.probs,
               sigma_states=inp.sigma_states)
    with tempfile.TemporaryDirectory(dir=workdir) as tmp:
        tmp = Path(tmp)
        pickle.dump({"object": obj, "alpha": float(alpha), "repeats": int(repeats)}, open(tmp / "in.pkl", "wb"))
        env = reference_env(tmp / "stubs", with_backend=False)
        out = subprocess.run([sys.executable, "-c", TIME_CALC_VAR, str(tmp / "in.pkl"), str(tmp / "out.pkl")], env=env,
                             cwd=tmp, capture_output=True, text=True, timeout=timeout)
        if out.returncode != 0:
            raise RuntimeError("reference run failed: " + out.stderr[-1500:])
        return pickle.load(open(tmp / "out.pkl", "rb"))
