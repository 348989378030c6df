"""Host-side driver of the CUDA VaR solve (ctypes over include/cvar.h).

`VarPlan` owns one `cvar_plan_t` (device copies of the axis, the Student-t quantile table, a stream and a
workspace).  Two ways in:

* host buffers (`strip_mass`, `solve`): NumPy in, NumPy out, H2D / kernels / D2H inside the C call --
  this is what the drop-in `ValueAtRiskCalcualtion.calc_var` uses;
* device buffers (`solve_device`, `finalize_device`, `strip_mass_device`): torch CUDA tensors, work is
  enqueued on torch's current stream and nothing synchronises -- used by the benchmark's device-resident
  leg and by the multi-GPU driver.

No CPU fallback exists: constructing a plan without the built library or without a CUDA device raises.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib
from .inputs import HotPathInputs


@dataclass
class SolveResult:
    var: np.ndarray          # (n_alpha, T) solved quantile + ptf_mean
    case: np.ndarray         # (n_alpha, T) bracket id 0..3 (A..D), 4 = undefined
    cells: np.ndarray        # (n_alpha, T) grid cells evaluated per solve
    iterations: np.ndarray   # (n_alpha,) global bisection iteration count K
    kernel_ms: float         # device time of solve + finalize kernels
    status: np.ndarray = None  # (n_alpha,) CVAR_STATUS_* bits of the finalize (zero-mass exit taken / ambiguous)


def _ptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def _f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None and a.shape != shape:
        raise ValueError(f"expected shape {shape}, got {a.shape}")
    return a


class VarPlan:
    """Run-constant state of the solve for one (copula, marginal family, grid, weights) combination."""

    def __init__(self, inputs: HotPathInputs, device: int | None = None, compat_flags: int = _lib.COMPAT_REFERENCE,
                 max_iter: int = 0, first_guess: float = -3.0, second_guess=(-3.5, -2.0), clip_lo: float = -5.0):
        lib = _lib.load()
        self._lib = lib
        d = _lib.CvarDesc()
        lib.cvar_desc_default(C.byref(d))
        d.copula = _lib.COPULA_ID[inputs.copula]
        d.marginal = _lib.MARGINAL_ID[inputs.marginal]
        d.n = int(inputs.n)
        d.q = int(inputs.q)
        d.compat_flags = int(compat_flags)
        d.max_iter = int(max_iter)
        d.rho, d.nu, d.theta = float(inputs.rho), float(inputs.nu), float(inputs.theta)
        d.w0, d.w1 = float(inputs.weights[0]), float(inputs.weights[1])
        d.first_guess = float(first_guess)
        d.second_lo, d.second_hi = float(second_guess[0]), float(second_guess[1])
        d.clip_lo = float(clip_lo)   # lower_bound of the reference's grid (calc_var_class.py:201)
        self.desc = d
        self.copula, self.marginal, self.n, self.q = inputs.copula, inputs.marginal, int(inputs.n), int(inputs.q)
        self._x = _f64(inputs.x, (self.n,))
        self._dx = _f64(inputs.dx, (self.n,))
        self._states = None if inputs.marginal == "single" else _f64(inputs.sigma_states, (2, self.q))
        handle = C.c_void_p()
        st = lib.cvar_plan_create(C.byref(d), _ptr(self._x), _ptr(self._dx), _ptr(self._states),
                                  -1 if device is None else int(device), C.byref(handle))
        _lib.check(st, "cvar_plan_create")
        self._h = handle
        self._reserved = 0
        self.device = self.info().device

    # -- lifetime ---------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self._lib.cvar_plan_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def reserve(self, days: int) -> int:
        """Size the plan's per-chunk device scratch for batches of `days` days (cvar_plan_reserve); returns the chunk size.

        The C `*_device` entry points never allocate: a batch larger than the reserved chunk runs in several chunks.
        `solve_device` / `strip_mass_device` below reserve for their batch on first use, so that a device-resident
        batch normally runs as one chunk."""
        if days > self._reserved:
            _lib.check(self._lib.cvar_plan_reserve(self._h, int(days)), "cvar_plan_reserve")
            self._reserved = max(int(days), int(self.info().chunk_days))
        return int(self.info().chunk_days)

    def info(self) -> _lib.CvarPlanInfo:
        info = _lib.CvarPlanInfo()
        _lib.check(self._lib.cvar_plan_get_info(self._h, C.byref(info)), "cvar_plan_get_info")
        return info

    @property
    def max_iter(self) -> int:
        return int(self.info().max_iter)

    def _day_shape(self, T):
        return (T, 2) if self.marginal == "single" else (T, 2, self.q)

    # -- host-buffer path -------------------------------------------------------------------
    def strip_mass(self, day_params, bounds, return_cells: bool = False):
        """Strip mass S(lo, hi) per day: the `compute_integral` seam (calc_var_class.py:179-212)."""
        bounds = _f64(bounds)
        T = bounds.shape[0]
        day = _f64(day_params, self._day_shape(T))
        out = np.empty(T)
        cells = np.empty(T, dtype=np.uint64) if return_cells else None
        st = self._lib.cvar_strip_mass_host(self._h, _ptr(day), T, _ptr(bounds), _ptr(out), _ptr(cells))
        _lib.check(st, "cvar_strip_mass_host")
        return (out, cells) if return_cells else out

    def solve(self, day_params, alphas, ptf_mean: float = 0.0, forced_iterations=None, out=None,
              details: bool = True) -> SolveResult:
        """All (day, alpha) solves of a batch: `calc_var` for every alpha (calc_var_class.py:95-177).

        ``out`` may be a preallocated (n_alpha, T) float64 array (e.g. pinned memory) receiving the VaR levels.
        ``details=False`` skips the bracket ids, cell counters, status words and kernel time (those fields of the result are
        None / NaN): nothing but the VaR vector and the iteration counts is copied back.
        """
        alphas = _f64(np.atleast_1d(alphas))
        na = alphas.shape[0]
        day = np.ascontiguousarray(day_params, dtype=np.float64)
        T = day.shape[0]
        if day.shape != self._day_shape(T):
            raise ValueError(f"day_params must have shape {self._day_shape(T)}, got {day.shape}")
        var = out if out is not None else np.empty((na, T))
        if var.shape != (na, T) or var.dtype != np.float64 or not var.flags.c_contiguous:
            raise ValueError("out must be a C-contiguous float64 array of shape (n_alpha, T)")
        case = np.empty((na, T), dtype=np.int32) if details else None
        cells = np.empty((na, T), dtype=np.uint64) if details else None
        iters = np.empty(na, dtype=np.int32)
        forced = None if forced_iterations is None else np.ascontiguousarray(
            np.broadcast_to(np.asarray(forced_iterations, dtype=np.int32), (na,)))
        st = self._lib.cvar_solve_host(self._h, _ptr(day), T, _ptr(alphas), na, _ptr(forced), float(ptf_mean),
                                       _ptr(var), _ptr(case), _ptr(cells), _ptr(iters))
        _lib.check(st, "cvar_solve_host")
        if not details:   # the lean path: nothing beyond the VaR levels and the iteration counts is fetched
            return SolveResult(var=var, case=None, cells=None, iterations=iters, kernel_ms=float("nan"), status=None)
        return SolveResult(var=var, case=case, cells=cells, iterations=iters, kernel_ms=float(self.info().last_kernel_ms),
                           status=self.last_status(na))

    def evaluated_cells(self, reset: bool = True) -> int:
        """Cells the plan's solve launches really evaluated since the last reset (cvar_evaluated_cells_host): the
        numerator of the issued-instruction roofline; strips shared between the alphas of a day count once."""
        total = C.c_uint64()
        _lib.check(self._lib.cvar_evaluated_cells_host(self._h, C.byref(total), int(bool(reset))), "cvar_evaluated_cells_host")
        return int(total.value)

    def last_status(self, n_alpha: int = 1) -> np.ndarray:
        """CVAR_STATUS_* words of the plan's last finalize, one per alpha (cvar_finalize_status_host)."""
        out = np.zeros(int(n_alpha), dtype=np.int32)
        _lib.check(self._lib.cvar_finalize_status_host(self._h, _ptr(out), int(n_alpha)), "cvar_finalize_status_host")
        return out

    def special(self, which: int, values) -> np.ndarray:
        """Device special functions (tests): see cvar_test_special_host."""
        v = _f64(np.atleast_1d(values))
        out = np.empty_like(v)
        _lib.check(self._lib.cvar_test_special_host(self._h, int(which), _ptr(v), v.size, _ptr(out)),
                   "cvar_test_special_host")
        return out

    # -- device-buffer path (torch tensors) -----------------------------------------------------
    def _check_tensor(self, t, dtype, name):
        import torch

        if not (isinstance(t, torch.Tensor) and t.is_cuda and t.is_contiguous() and t.dtype == dtype):
            raise ValueError(f"{name} must be a contiguous CUDA tensor of dtype {dtype}")
        if t.device.index != self.device:
            raise ValueError(f"{name} lives on cuda:{t.device.index}, the plan on cuda:{self.device}")

    def solve_device(self, day_params, alphas, traj=None, mass=None, cells=None, reserve: bool = True):
        """Enqueue the solve kernel on torch's current stream. Returns the trajectory tensor (n_alpha, T, 2) int32.

        ``reserve=False`` leaves the plan's chunk size as it is (a larger batch then runs in several chunks)."""
        import torch

        self._check_tensor(day_params, torch.float64, "day_params")
        T = day_params.shape[0]
        alphas = _f64(np.atleast_1d(alphas))
        na = alphas.shape[0]
        if traj is None:
            traj = torch.empty((na, T, 2), dtype=torch.int32, device=day_params.device)
        self._check_tensor(traj, torch.int32, "traj")
        if mass is not None:
            self._check_tensor(mass, torch.float64, "mass")
        if cells is not None:
            self._check_tensor(cells, torch.int64, "cells")
        if reserve:
            self.reserve(T)
        stream = torch.cuda.current_stream(day_params.device).cuda_stream
        st = self._lib.cvar_solve_device(self._h, C.c_void_p(day_params.data_ptr()), T, _ptr(alphas), na,
                                         C.c_void_p(traj.data_ptr()),
                                         None if mass is None else C.c_void_p(mass.data_ptr()),
                                         None if cells is None else C.c_void_p(cells.data_ptr()),
                                         C.c_void_p(stream))
        _lib.check(st, "cvar_solve_device")
        return traj

    def finalize_device(self, traj, ptf_mean: float = 0.0, forced_iterations=None, var=None, case=None, iterations=None,
                        T: int | None = None):
        """Enqueue the finalize kernels over a (possibly gathered) trajectory tensor.

        traj: (n_alpha, T, 2), or -- straight out of an all-gather, no reshuffling -- the blocked layout
        (n_blocks, n_alpha, block_days, 2) with day d in block d // block_days; pass the true number of days `T` when
        the last block is ragged (cvar_finalize_blocked_device)."""
        import torch

        self._check_tensor(traj, torch.int32, "traj")
        if traj.dim() == 4:
            nb, na, block = traj.shape[0], traj.shape[1], traj.shape[2]
            T = nb * block if T is None else int(T)
            if not (nb - 1) * block < T <= nb * block and not (T == 0 and nb * block == 0):
                raise ValueError(f"T = {T} does not fit {nb} blocks of {block} days")
        else:
            na, block = traj.shape[0], max(traj.shape[1], 1)
            T = traj.shape[1]
        if var is None:
            var = torch.empty((na, T), dtype=torch.float64, device=traj.device)
        if case is None:
            case = torch.empty((na, T), dtype=torch.int32, device=traj.device)
        if iterations is None:
            iterations = torch.empty((na,), dtype=torch.int32, device=traj.device)
        forced = None if forced_iterations is None else np.ascontiguousarray(
            np.broadcast_to(np.asarray(forced_iterations, dtype=np.int32), (na,)))
        stream = torch.cuda.current_stream(traj.device).cuda_stream
        st = self._lib.cvar_finalize_blocked_device(self._h, C.c_void_p(traj.data_ptr()), T, block, na, _ptr(forced),
                                                    float(ptf_mean), C.c_void_p(var.data_ptr()), C.c_void_p(case.data_ptr()),
                                                    C.c_void_p(iterations.data_ptr()), C.c_void_p(stream))
        _lib.check(st, "cvar_finalize_blocked_device")
        return var, case, iterations

    def strip_mass_device(self, day_params, bounds, out=None):
        import torch

        self._check_tensor(day_params, torch.float64, "day_params")
        self._check_tensor(bounds, torch.float64, "bounds")
        T = bounds.shape[0]
        if out is None:
            out = torch.empty((T,), dtype=torch.float64, device=bounds.device)
        self.reserve(T)
        stream = torch.cuda.current_stream(bounds.device).cuda_stream
        st = self._lib.cvar_strip_mass_device(self._h, C.c_void_p(day_params.data_ptr()), T, C.c_void_p(bounds.data_ptr()),
                                              C.c_void_p(out.data_ptr()), None, C.c_void_p(stream))
        _lib.check(st, "cvar_strip_mass_device")
        return out


def fp64_peak_tflops(device: int = -1, min_ms: float = 50.0) -> tuple[float, float]:
    """(TFLOP/s, ms): FP64-pipe peak measured with a dependency-free DFMA micro-benchmark."""
    t, ms = C.c_double(), C.c_double()
    _lib.check(_lib.load().cvar_fp64_peak_host(int(device), float(min_ms), C.byref(t), C.byref(ms)), "cvar_fp64_peak_host")
    return t.value, ms.value


def solve_var(inputs: HotPathInputs, alphas, device: int | None = None, **plan_kw) -> SolveResult:
    """One-shot helper: build a plan for ``inputs`` and solve every (day, alpha)."""
    with VarPlan(inputs, device=device, **plan_kw) as plan:
        return plan.solve(inputs.day_params(), alphas, ptf_mean=inputs.ptf_mean)
