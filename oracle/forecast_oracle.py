"""CPU ORACLE of the forecast producers  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE (see var_oracle.py header).

Restates, in NumPy, the reference's rolling-window MSM state filter and GARCH one-step volatility forecast:
  * markov_switching_multifractal/calc_prob.py:8-69 (`calc_state_prob_numba`, `calc_bayes_upd_numba`),
    :86-122 (state table, transition matrix, state likelihoods), calc_marginals.py:33-38 (`calc_forecasts`)
  * garch/estimation.py:40-65 (`calculate_conditional_variances`), garch/forecast.py:5-18 (`calc_forecast`)
PARITY PIN: tests/golden/forecast_producers.npz holds the outputs of the unmodified reference functions
(tests/golden/make_golden_forecast.py); tests/test_forecast_oracle.py checks this module against them.
"""
from __future__ import annotations

import itertools

import numpy as np


def msm_tables(k, m0, sigma_bar, b, gamma):
    """(vol_states[2^k], transition matrix[2^k, 2^k]) exactly as ProbEstimation builds them (calc_prob.py:86-108)."""
    table = np.array(list(itertools.product([m0, 2 - m0], repeat=k)))
    gamma_k = 1 - (1 - gamma) ** (b ** np.arange(table.shape[1]))
    p_values = 1 - gamma_k / 2
    q_values = 1 - p_values
    P = np.prod(np.where(table[:, None, :] == table[None, :, :], p_values, q_values), axis=2)
    vol = np.array([np.sqrt(np.prod(row)) * sigma_bar for row in table])
    return vol, P


def msm_filter_last(returns, vol_states, P):
    """Filtered state distribution after the last return of one window (calc_prob.py:8-32, 110-122).
    numba's np.sum / += loops accumulate left to right, hence the cumulative sums."""
    S = len(vol_states)
    lik = (1 / (vol_states[None, :] * np.sqrt(2 * np.pi))) * np.exp(-0.5 * (returns[:, None] / vol_states[None, :]) ** 2)
    prev = np.full(S, 1 / S)
    for i in range(len(returns)):
        pred = np.cumsum(P * prev[None, :], axis=1)[:, -1]
        prob = pred * lik[i]
        scale = np.cumsum(prob)[-1]
        if scale == 0:
            return np.full(S, -1.0)
        prev = prob / scale
    return prev


def msm_forecast(series, k, m0, sigma_bar, b, gamma, N, T):
    """state_probs[T, 2^k]: one filter run per rolling window series[t:t+N]."""
    vol, P = msm_tables(k, m0, sigma_bar, b, gamma)
    return np.array([msm_filter_last(np.asarray(series[t:t + N], float), vol, P) for t in range(T)])


def garch_forecast_one(omega, alpha_vect, beta_vect, returns):
    """sqrt of the one-step-ahead conditional variance (garch/forecast.py:5-18)."""
    alpha_vect, beta_vect = np.asarray(alpha_vect, float), np.asarray(beta_vect, float)
    p, q, n = len(alpha_vect), len(beta_vect), len(returns)
    sigma2 = np.zeros(n)
    sigma2[0] = omega / (1 - sum(alpha_vect) - sum(beta_vect))
    for t in range(1, n):
        s = omega
        for i in range(min(p, t)):
            s += alpha_vect[i] * (returns[t - i - 1] ** 2)
        for j in range(min(q, t)):
            s += beta_vect[j] * sigma2[t - j - 1]
        sigma2[t] = max(s, 1e-7)
    forecast = omega + np.sum(alpha_vect * returns[-p:] ** 2) + np.sum(beta_vect * sigma2[-q:])
    return np.sqrt(forecast)


def garch_forecast(series, omega, alpha_vect, beta_vect, N, T):
    return np.array([garch_forecast_one(omega, alpha_vect, beta_vect, np.asarray(series[t:t + N], float)) for t in range(T)])


def kalman_forecast_one(returns, a, l, q, alpha=1.6, beta=2.0, kappa=1.75):
    """exp(last predicted state mean) of the scalar unscented filter (kalman_mean_reverting/estimate.py:231-281 with
    init_log_vol = l, init_var = q as forecast.py:9 passes them).  Returns NaN where the reference returns its error tuple."""
    L = 2
    lam = (alpha ** 2) * (L + kappa) - L
    wm = np.full(2 * L + 1, 1 / (2 * (L + lam)))
    wc = wm.copy()
    wm[0] = lam / (L + lam)
    wc[0] = wm[0] + (1 - alpha ** 2 + beta)
    wm2 = np.full(L + 1, 1 / (2 * (L + lam)))
    wm2[0] = lam / (L + lam)
    phi = np.sqrt(L + lam)
    mean, var, x_mean = l, q, 0.0
    for t in range(len(returns)):
        d = var if var > 0 else var + 1e-8                      # custom_cholesky of diag(var, 1)
        sv = np.sqrt(d)
        x1 = np.array([mean, mean + phi * sv, mean, mean - phi * sv, mean])
        x2 = np.array([0.0, 0.0, phi, 0.0, -phi])
        X = a * (x1 - l) + l + q * x2
        x_mean = np.dot(X, wm)
        diff = X - x_mean
        P = np.dot(diff * wc, diff)
        sp = np.sqrt(P)
        Y = np.array([x_mean, x_mean + phi * sp, x_mean - phi * sp])
        eta = returns[t] / np.exp(Y)
        h = (1 / np.sqrt(2 * np.pi)) * np.exp(-0.5 * eta ** 2) * np.abs(eta)
        Z = np.sum(wm2 * h)
        if Z <= 0 or Z < 1e-10:
            return np.nan
        mean = np.sum((wm2 * Y * h) / Z)
        var = np.sum(wm2 * ((h / Z) * (Y - mean) ** 2))
    return np.exp(x_mean)


def kalman_forecast(series, a, l, q, N, T):
    return np.array([kalman_forecast_one(np.asarray(series[t:t + N], float), a, l, q) for t in range(T)])
