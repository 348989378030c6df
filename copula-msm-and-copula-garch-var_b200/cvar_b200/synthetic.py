"""Seeded synthetic per-day forecast parameters.

There is no network, so neither the reference's Yahoo download nor its
in-sample fits can run inside a benchmark; the hot path is fed with
forecast parameters of the right shape and range instead:

* ``garch_sigma_path``   – sqrt of a simulated GARCH(1,1) variance path
* ``kalman_sigma_path``  – exp of a simulated AR(1) log-vol state
* ``msm_day_params``     – binomial MSM(k): 2^k vol states, Hamilton-filtered
  state probabilities along a simulated return path, merged to the k+1
  distinct vol levels the way the reference's adapter merges them
  (utils/model_estimation/model/msm_estimation.py:205-248)

The MSM building blocks restate the model of the reference's
markov_switching_multifractal/calc_prob.py:72-108 (state enumeration in
`itertools.product` order, per-component switching probabilities
``gamma_i = 1-(1-gamma)**(b**i)``).
"""
from __future__ import annotations

import numpy as np

from .inputs import HotPathInputs, make_inputs
from .msm_layout import (merge_states, msm_multiplier_table, msm_switch_probs, msm_transition_matrix,  # noqa: F401
                         msm_vol_states)

# asset-level constants of SURVEY §8(d)
GARCH_ASSETS = ((0.02, 0.09, 0.89, 1), (0.03, 0.08, 0.90, 2))       # omega, alpha, beta, seed
KALMAN_ASSETS = ((0.97, 0.0, 0.15, 3), (0.97, 0.0, 0.15, 4))         # a, l, q, seed
MSM_ASSETS = ((0.4, 1.1, 3.0, 0.3, 5), (0.55, 1.4, 5.0, 0.2, 6))     # m0, sigma_bar, b, gamma, seed


def garch_sigma_path(T: int, assets=GARCH_ASSETS) -> np.ndarray:
    """sigma[T, 2]: volatility (percent units) along a simulated GARCH(1,1) path."""
    out = np.empty((T, len(assets)))
    for a, (omega, alpha, beta, seed) in enumerate(assets):
        rng = np.random.default_rng(seed)
        s2 = omega / (1.0 - alpha - beta)
        for t in range(T):
            if t > 0:
                e = rng.standard_normal()
                s2 = omega + alpha * s2 * e * e + beta * s2
            out[t, a] = np.sqrt(s2)
    return out


def kalman_sigma_path(T: int, assets=KALMAN_ASSETS) -> np.ndarray:
    """sigma[T, 2] = exp(X_t) with X an AR(1) log-vol state (mean-reverting model)."""
    out = np.empty((T, len(assets)))
    for a, (phi, level, q, seed) in enumerate(assets):
        rng = np.random.default_rng(seed)
        xs = level
        for t in range(T):
            if t > 0:
                xs = phi * (xs - level) + level + q * rng.standard_normal()
            out[t, a] = np.exp(xs)
    return out


# ----------------------------------------------------------------------------
# binomial MSM(k)
# ----------------------------------------------------------------------------
def msm_simulate_returns(T: int, k: int, m0: float, sigma_bar: float, b: float, gamma: float, seed: int):
    """One simulated MSM return path of length T (percent units)."""
    rng = np.random.default_rng(seed)
    g = msm_switch_probs(k, b, gamma)
    mult = np.where(rng.random(k) < 0.5, m0, 2.0 - m0)
    r = np.empty(T)
    for t in range(T):
        switch = rng.random(k) < g
        draw = np.where(rng.random(k) < 0.5, m0, 2.0 - m0)
        mult = np.where(switch, draw, mult)
        r[t] = sigma_bar * np.sqrt(np.prod(mult)) * rng.standard_normal()
    return r


def hamilton_filter(returns: np.ndarray, vol_states: np.ndarray, P: np.ndarray) -> np.ndarray:
    """Filtered state probabilities pi[t, s] = P(state_t = s | r_1..r_t).

    Uniform prior, predict with ``P @ pi`` then Bayes update with the normal
    likelihood of r_t under each state's volatility (the recursion of the
    reference's markov_switching_multifractal/calc_prob.py:8-69).
    """
    T = len(returns)
    S = len(vol_states)
    lik = np.exp(-0.5 * (returns[:, None] / vol_states[None, :]) ** 2) / (vol_states[None, :] * np.sqrt(2 * np.pi))
    pi = np.full(S, 1.0 / S)
    out = np.empty((T, S))
    for t in range(T):
        pred = P @ pi
        post = pred * lik[t]
        s = post.sum()
        if s == 0.0:
            raise FloatingPointError("Hamilton filter degenerated (zero scaling factor)")
        pi = post / s
        out[t] = pi
    return out


def msm_day_params(T: int, k: int = 8, assets=MSM_ASSETS):
    """(probs_by_state[T,2,k+1], sigma_states[2,k+1]) for the synthetic MSM run."""
    vols, probs = [], []
    for (m0, sbar, b, gamma, seed) in assets:
        v = msm_vol_states(k, m0, sbar)
        P = msm_transition_matrix(k, m0, b, gamma)
        r = msm_simulate_returns(T, k, m0, sbar, b, gamma, seed)
        vols.append(v)
        probs.append(hamilton_filter(r, v, P))
    return merge_states(np.array(vols), np.array(probs))


# ----------------------------------------------------------------------------
# BASELINE.json configurations
# ----------------------------------------------------------------------------
def baseline_config(name: str, T: int | None = None, n: int | None = None) -> tuple[HotPathInputs, tuple[float, ...]]:
    """(inputs, alphas) of one BASELINE.json configuration; T / n may be overridden.

    c1: Gaussian + GARCH sigma path, n=100,  T=250,  99 % VaR
    c2: Student-t + GARCH,           n=1024, T=1000, 95 % / 99 %
    c3: Student-t + MSM k=8 (q=9),   n=2048, T=1000, 99 %
    c4: Plackett + Kalman sigma,     n=2048, T=1000, 95 % / 99 %
    c5_<copula>_<marginal>: throughput sweep member, n=4096, two alphas
    """
    if name == "c1":
        T, n = T or 250, n or 100
        return make_inputs("gaussian", "single", n, rho=0.6, sigma=garch_sigma_path(T)), (0.01,)
    if name == "c2":
        T, n = T or 1000, n or 1024
        return make_inputs("student", "single", n, rho=0.6, nu=5.3, sigma=garch_sigma_path(T)), (0.05, 0.01)
    if name == "c3":
        T, n = T or 1000, n or 2048
        pbs, lv = msm_day_params(T, 8)
        return make_inputs("student", "mixture", n, rho=0.6, nu=5.3, probs=pbs, sigma_states=lv), (0.01,)
    if name == "c4":
        T, n = T or 1000, n or 2048
        return make_inputs("plackett", "single", n, theta=4.2, sigma=kalman_sigma_path(T)), (0.05, 0.01)
    if name.startswith("c5_"):
        _, copula, marginal = name.split("_")
        T, n = T or 100_000, n or 4096
        if marginal == "single":
            return make_inputs(copula, "single", n, sigma=garch_sigma_path(T)), (0.01, 0.05)
        pbs, lv = msm_day_params(T, 8)
        return make_inputs(copula, "mixture", n, probs=pbs, sigma_states=lv), (0.01, 0.05)
    raise ValueError(f"unknown configuration {name!r}")
