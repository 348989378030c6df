"""Elementwise GPU evaluations behind the calculators' `copula_density` / `integrated_function` hooks.

These exist so that the plugin API of the reference is complete; the VaR solve itself never goes through
them (it never materialises point lists).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


def copula_density_gpu(copula: str, cdf, *, rho=float("nan"), nu=float("nan"), theta=float("nan"), device: int = -1):
    """c(u) for an (P, 2) array of PIT values, evaluated on the GPU."""
    u = np.ascontiguousarray(cdf, dtype=np.float64)
    if u.ndim != 2 or u.shape[1] != 2:
        raise ValueError("cdf must have shape (P, 2): the GPU backend covers two-asset portfolios")
    out = np.empty(u.shape[0])
    st = _lib.load().cvar_copula_density_host(_lib.COPULA_ID[copula], float(rho), float(nu), float(theta),
                                              C.c_void_p(u.ctypes.data), u.shape[0], C.c_void_p(out.ctypes.data), device)
    _lib.check(st, "cvar_copula_density_host")
    return out


def integrand_single_gpu(grids, step_sizes, sigma, nu, corr_matrix, copula: str | None = None, theta=None):
    """Single-normal integrand on an explicit point list, (P, 1) like the reference's
    integration_functions/garch_integration_function.py:5-52 (copula density on the GPU, the per-point
    normal cdf/pdf scaling is cheap elementwise NumPy)."""
    from scipy.special import erf

    grids = np.asarray(grids, float)
    sigma = np.asarray(sigma, float)
    z = grids / sigma
    cdf = 0.5 * (1 + erf(z / np.sqrt(2)))
    pdf = (1 / np.sqrt(2 * np.pi)) * np.exp(-0.5 * z ** 2) / sigma
    if corr_matrix is None:
        c = copula_density_gpu("plackett", cdf, theta=float(nu))
    elif nu is None:
        c = copula_density_gpu("gaussian", cdf, rho=float(np.asarray(corr_matrix)[0, 1]))
    else:
        c = copula_density_gpu("student", cdf, nu=float(nu), rho=float(np.asarray(corr_matrix)[0, 1]))
    val = np.nan_to_num(c * np.prod(pdf, axis=1)).reshape(-1, 1)
    return val * np.asarray(step_sizes, float)
