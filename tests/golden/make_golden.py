"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py [--only NAME_SUBSTRING]

For every case the per-day inputs are produced by this repo's seeded
generators (cvar_b200.synthetic / explicit formulas of SURVEY App. C), handed
to the reference through ``_ref_worker.py`` (a subprocess with
``PYTHONPATH=<stubs>:/root/reference``), and the reference's outputs
(`compute_integral` strip masses, `calc_var` vectors, its axis arrays, its
state merge) are stored next to the inputs in ``tests/golden/<case>.npz``.
The reference has no tests of its own, so these files are the parity pin.
"""
from __future__ import annotations

import argparse
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
from cvar_b200.inputs import make_inputs                    # noqa: E402

REFERENCE = Path("/root/reference")


def _bounds(T, seed):
    rng = np.random.default_rng(seed)
    lo = rng.uniform(-7.5, -0.2, T)
    hi = lo + rng.uniform(0.0005, 2.0, T)
    b = np.column_stack([lo, np.minimum(hi, 0.5)])
    b[::3, 0] = -100.0                                       # some half-plane probes
    return b


def _kat1_mixture():
    k = 2
    vs = np.array([syn.msm_vol_states(k, 0.4, 1.1), syn.msm_vol_states(k, 0.55, 1.4)])
    T = 4
    probs = np.empty((2, T, 4))
    for a in range(2):
        for t in range(T):
            for j in range(4):
                probs[a, t, j] = 1 + ((t + 1) * (j + 1) * (a + 2)) % 5
    probs /= probs.sum(axis=2, keepdims=True)
    return vs, probs


def _msm8(T):
    vols, probs = [], []
    for (m0, sbar, b, gamma, seed) in syn.MSM_ASSETS:
        v = syn.msm_vol_states(8, m0, sbar)
        P = syn.msm_transition_matrix(8, m0, b, gamma)
        r = syn.msm_simulate_returns(T, 8, m0, sbar, b, gamma, seed)
        vols.append(v)
        probs.append(syn.hamilton_filter(r, v, P))
    return np.array(vols), np.array(probs)


def build_cases():
    cases = []

    def add(name, copula, marginal, n, *, est=None, alphas=(0.01, 0.05), weights=(0.5, 0.5), ptf_mean=0.0,
            sigma=None, vs_probs=None, strips=True, **cop):
        kw = {}
        merge = None
        if marginal == "single":
            kw["sigma"] = np.asarray(sigma, float)
        else:
            vs, pr = vs_probs
            pbs, lv = syn.merge_states(vs, pr)
            kw["probs"], kw["sigma_states"] = pbs, lv
            merge = (vs, pr)
        inp = make_inputs(copula, marginal, n, weights=weights, ptf_mean=ptf_mean, **kw, **cop)
        T = inp.T
        case = dict(name=name, copula_type=copula, marginal=marginal, n=n, weights=np.asarray(weights, float),
                    estimation=est or ("garch" if marginal == "single" else "msm"), ptf_mean=ptf_mean,
                    copula_params=inp.copula_params(), alphas=tuple(alphas), merge=merge,
                    bounds=_bounds(T, len(cases) + 11) if strips else None,
                    rho=inp.rho, nu=inp.nu, theta=inp.theta)
        if marginal == "single":
            case["sigma"] = inp.sigma
        else:
            case["probs"], case["sigma_states"] = inp.probs, inp.sigma_states
        cases.append(case)

    # --- KAT set 1: n=64, T=4 ---------------------------------------
    sig_kat1 = np.array([[0.8 + 0.3 * t, 0.9 + 0.25 * t] for t in range(4)])
    for cop in ("gaussian", "student", "plackett"):
        add(f"kat1_{cop}_single", cop, "single", 64, sigma=sig_kat1)
        add(f"kat1_{cop}_mixture", cop, "mixture", 64, vs_probs=_kat1_mixture())
    # --- KAT set 2: Gaussian n=100 T=6 ----------------------------------------------
    sig_kat2 = np.column_stack([np.linspace(0.6, 2.5, 6), np.linspace(0.7, 2.2, 6)])
    add("kat2_gaussian_n100", "gaussian", "single", 100, sigma=sig_kat2, alphas=(0.05,))
    cases[-1]["bounds"] = np.array([(-100, -3), (-3, -2), (-2.5, -2.25), (-1.7, -1.69), (-3.5, -3), (-100, -0.3)], float)
    # --- BASELINE config 1 in full ---------------------------------------------------
    add("c1_gaussian_garch_n100_T250", "gaussian", "single", 100, sigma=syn.garch_sigma_path(250), strips=False)
    # --- Student-t + single normal, n=100 --------------------------------------------
    add("student_garch_n100_T24", "student", "single", 100, sigma=syn.garch_sigma_path(24))
    # --- Plackett + Kalman sigma, unequal weights, non-zero portfolio mean -----------
    add("plackett_mr_w37_n80", "plackett", "single", 80, est="mean_reverting", weights=(0.3, 0.7), ptf_mean=0.0123,
        sigma=syn.kalman_sigma_path(5))
    add("gaussian_garch_w73_n72", "gaussian", "single", 72, weights=(0.7, 0.3), sigma=syn.garch_sigma_path(5))
    # --- MSM k=8 -> q=9 merged mixture ------------------------------------------------
    vp8 = _msm8(3)
    for cop in ("gaussian", "student", "plackett"):
        add(f"{cop}_msm8_n48", cop, "mixture", 48, vs_probs=vp8)
    add("student_msm8_w46_n40", "student", "mixture", 40, vs_probs=vp8, weights=(0.4, 0.6), alphas=(0.01,))
    # --- parameter sweeps -------------------------------------------------------------
    for rho in (-0.5, 0.3, 0.9):
        add(f"sweep_gaussian_rho{rho}", "gaussian", "single", 64, sigma=sig_kat1, rho=rho)
    for nu in (2.5, 8.11, 30.0):
        add(f"sweep_student_nu{nu}", "student", "single", 64, sigma=sig_kat1, nu=nu, rho=0.6)
    add("sweep_student_nu4_rho-0.4", "student", "single", 64, sigma=sig_kat1, nu=4.0, rho=-0.4)
    for theta in (0.5, 2.0, 20.0):
        add(f"sweep_plackett_theta{theta}", "plackett", "single", 64, sigma=sig_kat1, theta=theta)
    # --- bracket cases A..D and saturated erf (u == 0 / 1 -> NaN -> 0, quirks Q5/Q14) --
    sig_cases = np.array([[0.45, 0.5], [0.5, 0.3], [0.58, 0.57], [1.0, 1.2], [1.6, 1.8], [2.2, 2.4], [3.0, 2.8], [4.0, 4.5]])
    for cop in ("gaussian", "student", "plackett"):
        add(f"cases_{cop}_single", cop, "single", 64, sigma=sig_cases)
    # --- larger grids (reference still tractable) --------------------------------------
    add("gaussian_garch_n512_T4", "gaussian", "single", 512, sigma=syn.garch_sigma_path(4), alphas=(0.01,))
    add("student_garch_n256_T3", "student", "single", 256, sigma=syn.garch_sigma_path(3), alphas=(0.01,))
    add("plackett_mr_n1024_T2", "plackett", "single", 1024, est="mean_reverting", sigma=syn.kalman_sigma_path(2),
        alphas=(0.05,), strips=False)
    add("student_garch_n512_T2", "student", "single", 512, sigma=syn.garch_sigma_path(2), alphas=(0.01, 0.05), nu=5.3, rho=0.6)
    add("gaussian_garch_n1024_T2", "gaussian", "single", 1024, sigma=syn.garch_sigma_path(2), alphas=(0.01,), strips=False)
    add("student_msm8_n96_T2", "student", "mixture", 96, vs_probs=_msm8(2), alphas=(0.01, 0.05))
    add("plackett_msm8_n96_w64", "plackett", "mixture", 96, vs_probs=_msm8(2), weights=(0.6, 0.4), alphas=(0.01,))
    # --- round-1 additions: odd / tiny grids, extreme parameters on the mixture path, other alphas -----------------
    add("x_student_msm8_rho-0.5_w37_n56", "student", "mixture", 56, vs_probs=vp8, weights=(0.3, 0.7), rho=-0.5, nu=3.2,
        alphas=(0.01, 0.05))
    add("x_gaussian_msm8_rho-0.7_n60", "gaussian", "mixture", 60, vs_probs=vp8, rho=-0.7, alphas=(0.05,))
    add("x_plackett_kat1mix_theta0.5_n52", "plackett", "mixture", 52, vs_probs=_kat1_mixture(), theta=0.5, alphas=(0.01, 0.05))
    add("x_student_nu2.01_n64", "student", "single", 64, sigma=sig_kat1, nu=2.01, rho=0.3, alphas=(0.01,))
    add("x_student_nu50_n64", "student", "single", 64, sigma=sig_kat1, nu=50.0, rho=0.8, alphas=(0.05,))
    add("x_gaussian_odd_n101", "gaussian", "single", 101, sigma=sig_kat2, alphas=(0.01, 0.1))
    add("x_plackett_tiny_n37", "plackett", "single", 37, sigma=sig_kat1, alphas=(0.05, 0.2), theta=9.0)
    add("x_student_garch_alpha0.001_n72", "student", "single", 72, sigma=syn.garch_sigma_path(6), alphas=(0.001, 0.25))
    return cases


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None)
    args = ap.parse_args()
    if not REFERENCE.exists():
        raise SystemExit("the reference is not mounted at /root/reference; golden vectors can only be regenerated "
                         "in the build container")
    cases = build_cases()
    if args.only:
        cases = [c for c in cases if args.only in c["name"]]
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
    for case in cases:
        out = results[case["name"]]
        blob = {
            "copula": case["copula_type"], "marginal": case["marginal"], "n": case["n"], "weights": case["weights"],
            "ptf_mean": case["ptf_mean"], "rho": case["rho"], "nu": case["nu"], "theta": case["theta"],
            "alphas": np.asarray(case["alphas"], float), "ref_x": out["x"], "ref_dx": out["dx"],
        }
        if case["marginal"] == "single":
            blob["sigma"] = case["sigma"]
        else:
            blob["probs"], blob["sigma_states"] = case["probs"], case["sigma_states"]
            blob["raw_vol_states"], blob["raw_probs"] = case["merge"]
            blob["ref_merge_probs"], blob["ref_merge_sigma_states"] = out["merge_probs"], out["merge_sigma_states"]
        if case["bounds"] is not None:
            blob["bounds"], blob["ref_strip_mass"] = case["bounds"], out["strip_mass"]
        for a in case["alphas"]:
            blob[f"ref_var_{a}"] = out[f"var_{a}"]
        np.savez_compressed(HERE / f"{case['name']}.npz", **blob)
        print("wrote", case["name"])


if __name__ == "__main__":
    main()
