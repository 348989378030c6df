"""Golden vectors for the rounding-dependent zero-mass exit, from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_ambiguous.py

The wide fuzz (tools/fuzz_parity.py) found three random configurations on which the oracle leaves the bisection at K = 0
(every day's running mass cancels exactly in ITS sums) while the CUDA kernel runs on to the quantile: seeds 2133, 20559,
20794 of tests/test_gpu_random.py::_random_case.  This script hands exactly those inputs to the reference's own calc_var
(same worker as make_golden.py) and stores what it returns, so that the question "what does the reference do there"
has a committed answer: tests/golden/ambiguous_exit_seeds.npz.
"""
from __future__ import annotations

import os
import pickle
import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REPO = HERE.parent.parent
sys.path[:0] = [str(REPO / "copula-msm-and-copula-garch-var_b200"), str(REPO / "tests")]

from test_gpu_random import _random_case                      # noqa: E402

REFERENCE = Path("/root/reference")
SEEDS = (2133, 20559, 20794)


def main():
    if not REFERENCE.exists():
        raise SystemExit("the reference is not mounted at /root/reference")
    cases, inputs = [], {}
    for seed in SEEDS:
        inp, alphas = _random_case(seed)
        assert inp.marginal == "mixture"
        inputs[seed] = (inp, alphas)
        cases.append(dict(name=f"seed{seed}", copula_type=inp.copula, marginal="mixture", n=inp.n,
                          weights=np.asarray(inp.weights, float), estimation="msm", ptf_mean=inp.ptf_mean,
                          copula_params=inp.copula_params(), alphas=tuple(alphas), merge=None, bounds=None,
                          probs=inp.probs, sigma_states=inp.sigma_states))
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
        subprocess.run([sys.executable, str(HERE / "_ref_worker.py"), str(fin), str(fout)], check=True, env=env, cwd=tmp)
        with open(fout, "rb") as f:
            results = pickle.load(f)
    blob = {"seeds": np.asarray(SEEDS)}
    for seed, (inp, alphas) in inputs.items():
        out = results[f"seed{seed}"]
        assert np.array_equal(out["x"], inp.x)
        blob[f"{seed}_alphas"] = np.asarray(alphas, float)
        blob[f"{seed}_probs"] = inp.probs                       # guards the test against a drift of _random_case
        for k, a in enumerate(alphas):
            blob[f"{seed}_ref_var_{k}"] = out[f"var_{a}"]
    np.savez_compressed(HERE / "ambiguous_exit_seeds.npz", **blob)
    print("wrote ambiguous_exit_seeds.npz")


if __name__ == "__main__":
    main()
