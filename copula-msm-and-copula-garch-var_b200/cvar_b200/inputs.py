"""Plain-data description of one VaR run: everything the per-day solve reads.

This is the host-side picture of what the reference keeps spread over
``ValueAtRiskCalcualtion`` attributes (reference: utils/calc_var_class.py:23-45):
``copula_params``, ``integrations_params_t``, ``integrations_params_static``,
``grids_generations_params``, ``weights``, ``ptf_mean``.  Both the oracle and
the CUDA backend consume this one structure, so a parity test feeds the very
same arrays to both.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from .axis import build_axis

COPULAS = ("gaussian", "student", "plackett")
MARGINALS = ("single", "mixture")


@dataclass
class HotPathInputs:
    """Inputs of the per-day VaR solve (all float64, host memory).

    copula    : 'gaussian' | 'student' | 'plackett'
    marginal  : 'single' (one normal per asset per day: GARCH, Kalman) |
                'mixture' (q-state normal mixture per asset per day: MSM)
    n         : grid points per axis
    x, dx     : axis and right-endpoint step sizes, shape (n,)
    weights   : portfolio weights (w0, w1)
    rho,nu,theta : copula parameters (unused ones stay NaN)
    sigma     : (T, 2) per-day vol forecast            [single]
    probs     : (T, 2, q) per-day state probabilities  [mixture]
    sigma_states : (2, q) merged vol states            [mixture]
    ptf_mean  : scalar added to the solved quantile
    """

    copula: str
    marginal: str
    n: int
    x: np.ndarray
    dx: np.ndarray
    weights: np.ndarray = field(default_factory=lambda: np.array([0.5, 0.5]))
    rho: float = float("nan")
    nu: float = float("nan")
    theta: float = float("nan")
    sigma: np.ndarray | None = None
    probs: np.ndarray | None = None
    sigma_states: np.ndarray | None = None
    ptf_mean: float = 0.0

    def __post_init__(self):
        if self.copula not in COPULAS:
            raise ValueError(f"unknown copula {self.copula!r}")
        if self.marginal not in MARGINALS:
            raise ValueError(f"unknown marginal {self.marginal!r}")
        self.x = np.ascontiguousarray(self.x, dtype=np.float64)
        self.dx = np.ascontiguousarray(self.dx, dtype=np.float64)
        self.weights = np.ascontiguousarray(self.weights, dtype=np.float64)
        if self.x.shape != (self.n,) or self.dx.shape != (self.n,):
            raise ValueError("x/dx must have shape (n,)")
        if self.weights.shape != (2,):
            raise ValueError("only two-asset portfolios are on the hot path (dim == 2)")
        if self.marginal == "single":
            if self.sigma is None:
                raise ValueError("single-normal marginals need sigma[T,2]")
            self.sigma = np.ascontiguousarray(self.sigma, dtype=np.float64)
            if self.sigma.ndim != 2 or self.sigma.shape[1] != 2:
                raise ValueError("sigma must have shape (T, 2)")
        else:
            if self.probs is None or self.sigma_states is None:
                raise ValueError("mixture marginals need probs[T,2,q] and sigma_states[2,q]")
            self.probs = np.ascontiguousarray(self.probs, dtype=np.float64)
            self.sigma_states = np.ascontiguousarray(self.sigma_states, dtype=np.float64)
            if self.probs.ndim != 3 or self.probs.shape[1] != 2:
                raise ValueError("probs must have shape (T, 2, q)")
            if self.sigma_states.shape != (2, self.probs.shape[2]):
                raise ValueError("sigma_states must have shape (2, q)")

    # ------------------------------------------------------------------
    @property
    def T(self) -> int:
        return int(self.sigma.shape[0] if self.marginal == "single" else self.probs.shape[0])

    @property
    def q(self) -> int:
        return 1 if self.marginal == "single" else int(self.probs.shape[2])

    def day_params(self) -> np.ndarray:
        """The per-day parameter block as one contiguous array: (T,2) or (T,2,q)."""
        return self.sigma if self.marginal == "single" else self.probs

    def take_days(self, idx) -> "HotPathInputs":
        """Same run restricted to the days ``idx`` (slice or index array)."""
        kw = dict(self.__dict__)
        if self.marginal == "single":
            kw["sigma"] = self.sigma[idx]
        else:
            kw["probs"] = self.probs[idx]
        return HotPathInputs(**kw)

    def copula_params(self):
        """The reference's packed ``copula_params`` (student_estimation.py:23-38,
        gaussian_estimation.py:35-44, plackett_estimation.py:29-37)."""
        if self.copula == "gaussian":
            return np.array([self.rho])
        if self.copula == "student":
            return np.array([self.nu, self.rho])
        return float(self.theta)


def make_inputs(copula: str, marginal: str, n: int, *, weights=(0.5, 0.5), rho=0.6, nu=5.3,
                theta=4.2, sigma=None, probs=None, sigma_states=None, ptf_mean=0.0) -> HotPathInputs:
    """Convenience constructor that also builds the reference axis for ``n``."""
    x, dx = build_axis(n, marginal)
    kw = dict(copula=copula, marginal=marginal, n=int(n), x=x, dx=dx, weights=np.asarray(weights, float),
              sigma=sigma, probs=probs, sigma_states=sigma_states, ptf_mean=float(ptf_mean))
    if copula == "gaussian":
        kw["rho"] = float(rho)
    elif copula == "student":
        kw["rho"], kw["nu"] = float(rho), float(nu)
    else:
        kw["theta"] = float(theta)
    return HotPathInputs(**kw)
