"""Runs the UNMODIFIED reference hot path on prepared inputs (golden-vector generation only).

Executed by ``make_golden.py`` in a subprocess with
``PYTHONPATH=<stub dir>:/root/reference`` (the stubs provide empty ``yfinance``
and ``matplotlib`` packages, which the reference imports at module top but
does not need on this path).  Never imported by tests or product code, and
never run on the GPU box (``/root/reference`` does not exist there).

Recipe: SURVEY.md App. C -- build the calculator through the reference's own
factory, create ``ValueAtRiskCalcualtion`` without running its constructor
(which would download data and fit models) and inject the hot-path
attributes; then call the reference's ``compute_integral`` / ``calc_var``.
"""
import contextlib
import io
import pickle
import sys
import time

import numpy as np

from utils.factory import ValueAtRiskCalculationFactory          # reference module
from utils.calc_var_class import ValueAtRiskCalcualtion          # reference module


def build(case):
    est = case["estimation"]
    m = ValueAtRiskCalculationFactory.create_var_calculator(case["copula_type"], est)
    v = object.__new__(ValueAtRiskCalcualtion)
    n = case["n"]
    v.num_points = n
    v.weights = np.asarray(case["weights"], float)
    v.dim = 2
    v.ptf_mean = float(case.get("ptf_mean", 0.0))
    if case["marginal"] == "single":
        sigma = np.asarray(case["sigma"], float)
        v.out_sample_N = sigma.shape[0]
        dens, x, dx = m.compute_normal_densities(2, n)
        v.grids_generations_params = (dens, x, dx, np.zeros((1, 2)))
        v.integrations_params_t = [sigma]
        v.integrations_params_static = None
    else:
        fbs = np.asarray(case["probs"], float)             # (T, 2, q)
        uvs = np.asarray(case["sigma_states"], float)      # (2, q)
        v.out_sample_N = fbs.shape[0]
        dens, x, dx = m.compute_normal_densities(uvs, n)
        v.grids_generations_params = (dens, x, dx, m.create_vol_combinations(uvs))
        v.integrations_params_t = (fbs, m.compute_forecast_combinations(fbs))
        v.integrations_params_static = uvs
    v.copula_params = case["copula_params"]
    v.copula_function = m.copula_density
    v.unpack_copula_params = m.unpack_copula_params
    v.integrated_function = m.integrated_function
    return m, v, x, dx


def main(path_in, path_out):
    with open(path_in, "rb") as f:
        cases = pickle.load(f)
    results = {}
    for case in cases:
        t0 = time.time()
        sink = io.StringIO()
        with contextlib.redirect_stdout(sink):
            m, v, x, dx = build(case)
            out = {"x": x, "dx": dx}
            if case.get("bounds") is not None:
                out["strip_mass"] = np.asarray(v.compute_integral(np.asarray(case["bounds"], float)))
            for alpha in case.get("alphas", ()):
                out[f"var_{alpha}"] = np.asarray(v.calc_var(obj_var=alpha))
            if case.get("merge") is not None:
                vs, fa = case["merge"]
                fbs, uvs = m.sum_forecast_by_state(np.asarray(vs), np.asarray(fa))
                out["merge_probs"], out["merge_sigma_states"] = fbs, uvs
        results[case["name"]] = out
        print(f"[ref] {case['name']}: {time.time() - t0:.1f}s", file=sys.stderr, flush=True)
    with open(path_out, "wb") as f:
        pickle.dump(results, f)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
