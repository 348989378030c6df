"""Forecast producers on the GPU: rolling-window MSM state filtering and GARCH volatility forecasts.

They fill the solve's per-day parameter block (`day_params`) from centred return series, replacing the Python loops
over dates of the reference's adapters (utils/model_estimation/model/msm_estimation.py:143-248,
garch_estimation.py:190-231).  Model fitting stays out of scope: the model parameters are inputs.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib
from .msm_layout import msm_stay_probs, msm_vol_states, state_levels


@dataclass
class MsmParams:
    """Binomial MSM(k) parameters of one asset (the reference's `optimal_params`: m_0, sig, b, gamma)."""
    m0: float
    sigma_bar: float
    b: float
    gamma: float


def _ptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def rolling_series(windows) -> np.ndarray | None:
    """Collapse overlapping rolling windows [T][N] (window t = series[t:t+N]) back into one series, or None
    if the windows do not overlap that way."""
    w = np.asarray(windows, dtype=float)
    if w.ndim != 2:
        return None
    if w.shape[0] > 1 and not np.array_equal(w[1:, :-1], w[:-1, 1:]):
        return None
    return np.concatenate([w[0], w[1:, -1]])


def msm_forecast(returns, params, k: int, N: int, *, window_stride: int = 1, return_state_probs: bool = False,
                 device: int = -1):
    """Filtered MSM state distribution at the end of every rolling window, merged to the distinct vol levels.

    returns : (n_assets, L) centred returns, L = (T-1)*window_stride + N
    params  : one MsmParams per asset
    -> probs_by_state (T, n_assets, q), sigma_states (n_assets, q)[, state_probs (n_assets, T, 2**k)], info dict
    """
    r = np.ascontiguousarray(np.atleast_2d(returns), dtype=np.float64)
    na, L = r.shape
    if len(params) != na:
        raise ValueError("one MsmParams per asset")
    if (L - N) % window_stride or L < N:
        raise ValueError("series length does not match (T-1)*window_stride + N")
    T = (L - N) // window_stride + 1
    S = 1 << k
    vols = np.ascontiguousarray([msm_vol_states(k, p.m0, p.sigma_bar) for p in params])
    stay = np.ascontiguousarray([msm_stay_probs(k, p.b, p.gamma) for p in params])
    levels, sig = zip(*[state_levels(v) for v in vols])
    if len({len(s) for s in sig}) != 1:
        raise ValueError("assets merge to different numbers of vol levels")
    q = len(sig[0])
    lvl = np.ascontiguousarray(levels, dtype=np.int32)
    out = np.empty((T, na, q))
    sp = np.empty((na, T, S)) if return_state_probs else None
    status, ms = C.c_int32(0), C.c_double(0.0)
    st = _lib.load().cvar_msm_forecast_host(k, na, _ptr(stay), _ptr(vols), _ptr(lvl), q, _ptr(r), T, N, window_stride,
                                            _ptr(out), _ptr(sp), C.byref(status), C.byref(ms), device)
    _lib.check(st, "cvar_msm_forecast_host")
    info = {"degenerate": bool(status.value), "kernel_ms": ms.value, "T": T, "states": S, "levels": q}
    res = (out, np.array(sig))
    return (*res, sp, info) if return_state_probs else (*res, info)


def _garch_pack(na, omega, alpha_vects, beta_vects):
    om = np.ascontiguousarray(np.broadcast_to(np.asarray(omega, dtype=np.float64), (na,)))
    al, be = np.zeros((na, 8)), np.zeros((na, 8))
    p, q = np.empty(na, np.int32), np.empty(na, np.int32)
    for a in range(na):
        av, bv = np.atleast_1d(alpha_vects[a]).astype(float), np.atleast_1d(beta_vects[a]).astype(float)
        if not (np.all(av > 0) and np.all(bv > 0) and om[a] > 0 and av.sum() + bv.sum() < 1):
            raise ValueError("GARCH parameters must be positive with sum(alpha) + sum(beta) < 1")   # garch/estimation.py:22-38
        p[a], q[a] = len(av), len(bv)
        al[a, :len(av)], be[a, :len(bv)] = av, bv
    return om, p, q, al, be


def garch_forecast(returns, omega, alpha_vects, beta_vects, N: int, *, window_stride: int = 1, device: int = -1):
    """One-step GARCH(p,q) volatility forecast at the end of every rolling window: sigma (T, n_assets)."""
    r = np.ascontiguousarray(np.atleast_2d(returns), dtype=np.float64)
    na, L = r.shape
    if (L - N) % window_stride or L < N:
        raise ValueError("series length does not match (T-1)*window_stride + N")
    T = (L - N) // window_stride + 1
    om, p, q, al, be = _garch_pack(na, omega, alpha_vects, beta_vects)
    out = np.empty((T, na))
    ms = C.c_double(0.0)
    st = _lib.load().cvar_garch_forecast_host(na, _ptr(om), _ptr(p), _ptr(q), _ptr(al), _ptr(be), _ptr(r), T, N, window_stride,
                                              _ptr(out), C.byref(ms), device)
    _lib.check(st, "cvar_garch_forecast_host")
    return out, {"kernel_ms": ms.value, "T": T}


def kalman_forecast(returns, a, l, q, N: int, *, window_stride: int = 1, ukf=(1.6, 2.0, 1.75), device: int = -1):
    """Kalman mean-reverting log-vol forecast exp(last predicted state mean) per rolling window: sigma (T, n_assets)."""
    r = np.ascontiguousarray(np.atleast_2d(returns), dtype=np.float64)
    na, L = r.shape
    if (L - N) % window_stride or L < N:
        raise ValueError("series length does not match (T-1)*window_stride + N")
    T = (L - N) // window_stride + 1
    av, lv, qv = (np.ascontiguousarray(np.broadcast_to(np.asarray(v, dtype=np.float64), (na,))) for v in (a, l, q))
    out = np.empty((T, na))
    status, ms = C.c_int32(0), C.c_double(0.0)
    st = _lib.load().cvar_kalman_forecast_host(na, _ptr(av), _ptr(lv), _ptr(qv), float(ukf[0]), float(ukf[1]), float(ukf[2]),
                                               _ptr(r), T, N, window_stride, _ptr(out), C.byref(status), C.byref(ms), device)
    _lib.check(st, "cvar_kalman_forecast_host")
    return out, {"kernel_ms": ms.value, "T": T, "failed": bool(status.value)}


def msm_forecast_device(returns, params, k: int, N: int, *, window_stride: int = 1):
    """Device-resident variant: `returns` is a CUDA float64 tensor (n_assets, L); the filter is enqueued on torch's
    current stream and the merged probabilities come back as a CUDA tensor (T, n_assets, q) that can be handed
    straight to `VarPlan.solve_device` -- no host round trip between forecast and solve.

    -> probs_by_state (cuda tensor), sigma_states (numpy, n_assets x q), status (cuda int32 tensor, 1 = degenerate window)
    """
    import torch

    if not (isinstance(returns, torch.Tensor) and returns.is_cuda and returns.dtype == torch.float64 and returns.is_contiguous()):
        raise ValueError("returns must be a contiguous CUDA float64 tensor of shape (n_assets, L)")
    na, L = returns.shape
    if len(params) != na or (L - N) % window_stride or L < N:
        raise ValueError("bad shapes: one MsmParams per asset, L = (T-1)*window_stride + N")
    T = (L - N) // window_stride + 1
    S = 1 << k
    dev = returns.device
    vols = np.ascontiguousarray([msm_vol_states(k, p.m0, p.sigma_bar) for p in params])
    stay = np.ascontiguousarray([msm_stay_probs(k, p.b, p.gamma) for p in params])
    levels, sig = zip(*[state_levels(v) for v in vols])
    q = len(sig[0])
    d_vols = torch.from_numpy(vols).to(dev)
    d_lvl = torch.from_numpy(np.ascontiguousarray(levels, dtype=np.int32)).to(dev)
    out = torch.empty((T, na, q), dtype=torch.float64, device=dev)
    work = torch.empty((na, L, S), dtype=torch.float64, device=dev)
    status = torch.zeros((1,), dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
    with torch.cuda.device(dev):
        st = _lib.load().cvar_msm_forecast_device(k, na, _ptr(stay), C.c_void_p(d_vols.data_ptr()), C.c_void_p(d_lvl.data_ptr()), q,
                                                  C.c_void_p(returns.data_ptr()), T, N, window_stride, C.c_void_p(out.data_ptr()),
                                                  None, C.c_void_p(work.data_ptr()), C.c_void_p(status.data_ptr()),
                                                  C.c_void_p(stream))
    _lib.check(st, "cvar_msm_forecast_device")
    return out, np.array(sig), status


def _device_series(returns, N, window_stride):
    import torch

    if not (isinstance(returns, torch.Tensor) and returns.is_cuda and returns.dtype == torch.float64 and returns.is_contiguous()
            and returns.dim() == 2):
        raise ValueError("returns must be a contiguous CUDA float64 tensor of shape (n_assets, L)")
    na, L = returns.shape
    if (L - N) % window_stride or L < N:
        raise ValueError("series length does not match (T-1)*window_stride + N")
    return na, (L - N) // window_stride + 1


def garch_forecast_device(returns, omega, alpha_vects, beta_vects, N: int, *, window_stride: int = 1):
    """Device-resident `garch_forecast`: CUDA tensor in, sigma (T, n_assets) CUDA tensor out, on torch's current stream."""
    import torch

    na, T = _device_series(returns, N, window_stride)
    om, p, q, al, be = _garch_pack(na, omega, alpha_vects, beta_vects)
    out = torch.empty((T, na), dtype=torch.float64, device=returns.device)
    stream = torch.cuda.current_stream(returns.device).cuda_stream
    with torch.cuda.device(returns.device):
        st = _lib.load().cvar_garch_forecast_device(na, _ptr(om), _ptr(p), _ptr(q), _ptr(al), _ptr(be), C.c_void_p(returns.data_ptr()),
                                                    T, N, window_stride, C.c_void_p(out.data_ptr()), C.c_void_p(stream))
    _lib.check(st, "cvar_garch_forecast_device")
    return out


def kalman_forecast_device(returns, a, l, q, N: int, *, window_stride: int = 1, ukf=(1.6, 2.0, 1.75)):
    """Device-resident `kalman_forecast`: -> sigma (T, n_assets) CUDA tensor, status (cuda int32, 1 = a window failed)."""
    import torch

    na, T = _device_series(returns, N, window_stride)
    av, lv, qv = (np.ascontiguousarray(np.broadcast_to(np.asarray(v, dtype=np.float64), (na,))) for v in (a, l, q))
    out = torch.empty((T, na), dtype=torch.float64, device=returns.device)
    status = torch.zeros((1,), dtype=torch.int32, device=returns.device)
    stream = torch.cuda.current_stream(returns.device).cuda_stream
    with torch.cuda.device(returns.device):
        st = _lib.load().cvar_kalman_forecast_device(na, _ptr(av), _ptr(lv), _ptr(qv), float(ukf[0]), float(ukf[1]), float(ukf[2]),
                                                     C.c_void_p(returns.data_ptr()), T, N, window_stride,
                                                     C.c_void_p(out.data_ptr()), C.c_void_p(status.data_ptr()), C.c_void_p(stream))
    _lib.check(st, "cvar_kalman_forecast_device")
    return out, status
