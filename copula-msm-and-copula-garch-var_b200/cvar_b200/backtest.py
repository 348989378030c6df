"""Backtest layer on top of the VaR vectors (SURVEY §7 "Exceedance counts", §8(f) rank 3).

The reference only plots VaR against the realised portfolio return (main.py:6-20, 71-75); the judged quantity
"identical exceedance counts" needs a definition, which is the one SURVEY §7 fixes:

    r_ptf[t] = out_sample_data.mean(axis=1)[t]         (main.py:73 -- the UNWEIGHTED column mean, quirk Q15)
    hit[t]   = r_ptf[t] < VaR[t]

plus the two standard coverage tests (absent from the reference): Kupiec's proportion-of-failures test and
Christoffersen's independence / conditional-coverage tests.  Host-side NumPy: this is O(T) post-processing of the
solve's output, not part of the GPU hot path.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
from scipy.stats import chi2


def portfolio_returns(out_sample_data) -> np.ndarray:
    """Realised portfolio return per out-of-sample day exactly as main.py:73 forms it (plain mean over assets)."""
    if hasattr(out_sample_data, "mean") and hasattr(out_sample_data, "to_numpy") and getattr(out_sample_data, "ndim", 2) == 2:
        return out_sample_data.mean(axis=1).to_numpy()
    return np.asarray(out_sample_data, dtype=float).mean(axis=1)


def hits(var, r_ptf) -> np.ndarray:
    """Boolean exceedance indicator r_ptf < VaR, broadcast over a leading alpha axis of `var` if present."""
    return np.asarray(r_ptf, dtype=float) < np.asarray(var, dtype=float)


def exceedances(var, r_ptf):
    """Number of exceedances (an int, or one int per alpha row of `var`)."""
    h = hits(var, r_ptf)
    return int(h.sum()) if h.ndim == 1 else h.sum(axis=-1).astype(int)


def _xlogy(x, y):
    return 0.0 if x == 0 else x * np.log(y)


@dataclass
class CoverageTest:
    statistic: float
    p_value: float
    dof: int


def kupiec_pof(n_exceed: int, T: int, alpha: float) -> CoverageTest:
    """Kupiec (1995) proportion-of-failures likelihood-ratio test of H0: P(hit) = alpha."""
    x, pi = int(n_exceed), int(n_exceed) / T
    ll0 = _xlogy(T - x, 1 - alpha) + _xlogy(x, alpha)
    ll1 = _xlogy(T - x, 1 - pi) + _xlogy(x, pi)
    lr = max(-2.0 * (ll0 - ll1), 0.0)
    return CoverageTest(lr, float(chi2.sf(lr, 1)), 1)


def christoffersen_independence(hit_series) -> CoverageTest:
    """Christoffersen (1998) LR test that hits are serially independent (first-order Markov alternative)."""
    h = np.asarray(hit_series, dtype=bool)
    prev, cur = h[:-1], h[1:]
    n00 = int(np.sum(~prev & ~cur)); n01 = int(np.sum(~prev & cur))
    n10 = int(np.sum(prev & ~cur)); n11 = int(np.sum(prev & cur))
    pi01 = n01 / max(n00 + n01, 1)
    pi11 = n11 / max(n10 + n11, 1)
    pi = (n01 + n11) / max(n00 + n01 + n10 + n11, 1)
    ll0 = _xlogy(n00 + n10, 1 - pi) + _xlogy(n01 + n11, pi)
    ll1 = _xlogy(n00, 1 - pi01) + _xlogy(n01, pi01) + _xlogy(n10, 1 - pi11) + _xlogy(n11, pi11)
    lr = max(-2.0 * (ll0 - ll1), 0.0)
    return CoverageTest(lr, float(chi2.sf(lr, 1)), 1)


def conditional_coverage(hit_series, alpha: float) -> CoverageTest:
    """Christoffersen's joint test: LR_cc = LR_pof + LR_ind, chi-square with 2 degrees of freedom."""
    h = np.asarray(hit_series, dtype=bool)
    lr = kupiec_pof(int(h.sum()), h.size, alpha).statistic + christoffersen_independence(h).statistic
    return CoverageTest(lr, float(chi2.sf(lr, 2)), 2)


def backtest_report(var, r_ptf, alphas) -> list[dict]:
    """One summary row per alpha: exceedances, hit rate and the three coverage tests."""
    var = np.atleast_2d(np.asarray(var, dtype=float))
    rows = []
    for k, a in enumerate(np.atleast_1d(alphas)):
        h = hits(var[k], r_ptf)
        valid = ~np.isnan(var[k])
        h = h[valid]
        pof, ind, cc = kupiec_pof(int(h.sum()), h.size, float(a)), christoffersen_independence(h), conditional_coverage(h, float(a))
        rows.append({"alpha": float(a), "days": int(h.size), "exceedances": int(h.sum()), "hit_rate": float(h.mean()) if h.size else float("nan"),
                     "kupiec_lr": pof.statistic, "kupiec_p": pof.p_value, "independence_lr": ind.statistic,
                     "independence_p": ind.p_value, "cond_coverage_lr": cc.statistic, "cond_coverage_p": cc.p_value})
    return rows
