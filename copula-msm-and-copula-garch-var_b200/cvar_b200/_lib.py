"""ctypes binding of include/cvar.h.  Loading fails loudly: there is no CPU fallback."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

from .build import LIB_PATH

ABI_VERSION = 1
COPULA_ID = {"gaussian": 0, "student": 1, "plackett": 2}
MARGINAL_ID = {"single": 0, "mixture": 1}
COMPAT_SIGMA_SWAP, COMPAT_CASE_C_SIGN, COMPAT_NAN_TO_NUM, COMPAT_REFERENCE = 1, 2, 4, 7
MAX_ALPHA = 8
STATUS_ZERO_EXIT_TAKEN, STATUS_ZERO_EXIT_AMBIGUOUS = 1, 2
CASE_NAMES = ("A", "B", "C", "D", "undefined")


class CvarDesc(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32), ("abi_version", C.c_uint32),
        ("copula", C.c_int32), ("marginal", C.c_int32), ("n", C.c_int32), ("q", C.c_int32),
        ("compat_flags", C.c_uint32), ("max_iter", C.c_int32),
        ("rho", C.c_double), ("nu", C.c_double), ("theta", C.c_double),
        ("w0", C.c_double), ("w1", C.c_double),
        ("clip_lo", C.c_double), ("neg_inf", C.c_double),
        ("first_guess", C.c_double), ("second_lo", C.c_double), ("second_hi", C.c_double),
        ("min_var", C.c_double), ("max_var", C.c_double), ("tol", C.c_double),
    ]


class CvarPlanInfo(C.Structure):
    _fields_ = [
        ("device", C.c_int32), ("sm_count", C.c_int32), ("max_iter", C.c_int32), ("ctas_per_sm", C.c_int32),
        ("threads_per_cta", C.c_int32), ("smem_bytes_per_cta", C.c_int32),
        ("tq_table_max_rel_err", C.c_double), ("last_kernel_ms", C.c_double),
        ("kernel_variant", C.c_int32), ("cluster4_capacity", C.c_int32),
        ("pow_octaves", C.c_int32), ("chunk_days", C.c_int64),
    ]


class CvarError(RuntimeError):
    def __init__(self, status: int, where: str):
        self.status = status
        msg = load().cvar_strerror(status).decode()
        super().__init__(f"{where}: {msg} (status {status})")


_PD = C.POINTER(C.c_double)
_PROTOTYPES = {
    "cvar_abi_version": (C.c_int, []),
    "cvar_desc_default": (None, [C.POINTER(CvarDesc)]),
    "cvar_check_dim": (C.c_int, [C.c_int32]),
    "cvar_strerror": (C.c_char_p, [C.c_int]),
    "cvar_plan_create": (C.c_int, [C.POINTER(CvarDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]),
    "cvar_plan_destroy": (C.c_int, [C.c_void_p]),
    "cvar_plan_get_info": (C.c_int, [C.c_void_p, C.POINTER(CvarPlanInfo)]),
    "cvar_plan_reserve": (C.c_int, [C.c_void_p, C.c_int64]),
    "cvar_strip_mass_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cvar_strip_mass_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cvar_solve_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cvar_finalize_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cvar_evaluated_cells_host": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64), C.c_int]),
    "cvar_finalize_blocked_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cvar_finalize_status_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "cvar_finalize_status_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32]),
    "cvar_solve_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cvar_test_special_host": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p]),
    "cvar_msm_forecast_host": (C.c_int, [C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int64,
                                          C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_double), C.c_int]),
    "cvar_msm_forecast_device": (C.c_int, [C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int64,
                                            C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cvar_garch_forecast_host": (C.c_int, [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                            C.c_int64, C.c_int64, C.c_void_p, C.POINTER(C.c_double), C.c_int]),
    "cvar_kalman_forecast_host": (C.c_int, [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_void_p,
                                             C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_double), C.c_int]),
    "cvar_garch_forecast_device": (C.c_int, [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                              C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]),
    "cvar_kalman_forecast_device": (C.c_int, [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_double,
                                               C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cvar_fp64_peak_host": (C.c_int, [C.c_int, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "cvar_copula_density_host": (C.c_int, [C.c_int32, C.c_double, C.c_double, C.c_double, C.c_void_p, C.c_int64, C.c_void_p, C.c_int]),
}
EXPORTED_SYMBOLS = tuple(_PROTOTYPES)

_lib = None


def load(path: str | Path | None = None):
    """dlopen libcvar_b200.so and attach prototypes.  Raises if the library was not built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    import os
    p = Path(path) if path else Path(os.environ.get("CVAR_B200_LIB", LIB_PATH))    # env override: A/B-test another build
    if not p.exists():
        raise ImportError(
            f"{p} is missing: build it with `python __graft_entry__.py` (nvcc, sm_100a). "
            "The VaR backend has no CPU fallback.")
    lib = C.CDLL(str(p))
    for name, (res, args) in _PROTOTYPES.items():
        fn = getattr(lib, name)        # AttributeError if a declared symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.cvar_abi_version() != ABI_VERSION:
        raise ImportError(f"{p}: ABI version {lib.cvar_abi_version()} != expected {ABI_VERSION}")
    if path is None:
        _lib = lib
    return lib


def check_dim(dim: int):
    """NotImplementedError (with the library's own message) for portfolios the backend refuses: dim != 2."""
    status = load().cvar_check_dim(int(dim))
    if status != 0:
        raise NotImplementedError(f"dim = {dim}: {load().cvar_strerror(status).decode()} (status {status})")


def check(status: int, where: str):
    if status != 0:
        raise CvarError(status, where)
