"""Runs the UNMODIFIED reference forecast producers on prepared inputs (golden-vector generation only).
Executed by make_golden_forecast.py with PYTHONPATH=<stubs>:/root/reference; never imported by tests or product code."""
import pickle
import sys

import numpy as np

from garch.forecast import calc_forecast as garch_calc_forecast                          # reference
from kalman_mean_reverting.forecast import calc_forecast as kalman_calc_forecast         # reference
from markov_switching_multifractal.calc_marginals import calc_forecasts as msm_calc_forecasts   # reference


def main(path_in, path_out):
    with open(path_in, "rb") as f:
        cases = pickle.load(f)
    out = {}
    for c in cases:
        series, N, T = np.asarray(c["series"], float), c["N"], c["T"]
        if c["kind"] == "msm":
            res = np.array([msm_calc_forecasts(c["k"], c["m0"], c["sigma_bar"], c["b"], c["gamma"], series[t:t + N])
                            for t in range(T)])
        elif c["kind"] == "kalman":
            res = np.array([kalman_calc_forecast(series[t:t + N], c["a"], c["l"], c["q"]) for t in range(T)])
        else:
            res = np.array([garch_calc_forecast(c["omega"], np.asarray(c["alpha"], float), np.asarray(c["beta"], float),
                                                series[t:t + N]) for t in range(T)])
        out[c["name"]] = res
        print("[ref]", c["name"], res.shape, file=sys.stderr, flush=True)
    with open(path_out, "wb") as f:
        pickle.dump(out, f)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
