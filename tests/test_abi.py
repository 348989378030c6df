"""The C-ABI library loads without a GPU and exports exactly what include/cvar.h declares (no compute calls)."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

from conftest import REPO

HEADER = REPO / "include" / "cvar.h"


@pytest.fixture(scope="module")
def lib():
    from cvar_b200 import _lib
    from cvar_b200.build import build_library
    build_library()
    return _lib.load()


def declared_functions():
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(cvar_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_expected_entry_points():
    names = declared_functions()
    for must in ("cvar_plan_create", "cvar_plan_destroy", "cvar_strip_mass_host", "cvar_strip_mass_device",
                 "cvar_solve_device", "cvar_finalize_device", "cvar_solve_host", "cvar_copula_density_host",
                 "cvar_fp64_peak_host", "cvar_strerror", "cvar_desc_default", "cvar_abi_version"):
        assert must in names


def test_library_exports_every_declared_symbol(lib):
    from cvar_b200 import _lib
    for name in declared_functions():
        assert hasattr(lib, name), f"{name} declared in cvar.h but not exported"
    assert set(_lib.EXPORTED_SYMBOLS) == set(declared_functions())


def test_abi_version_and_desc_defaults(lib):
    from cvar_b200 import _lib
    assert lib.cvar_abi_version() == 1
    d = _lib.CvarDesc()
    lib.cvar_desc_default(C.byref(d))
    assert d.struct_size == C.sizeof(_lib.CvarDesc) == 136
    assert (d.n, d.q, d.compat_flags, d.max_iter) == (100, 1, 7, 0)
    assert (d.w0, d.w1, d.clip_lo, d.neg_inf) == (0.5, 0.5, -5.0, -100.0)
    assert (d.first_guess, d.second_lo, d.second_hi, d.min_var, d.max_var, d.tol) == (-3.0, -3.5, -2.0, -7.5, 0.0, 1e-6)


def test_strerror_covers_status_codes(lib):
    for code in range(-8, 1):
        msg = lib.cvar_strerror(code).decode()
        assert msg and msg != "unknown status"
    assert "CPU fallback" in lib.cvar_strerror(-6).decode()


def test_argument_validation_happens_before_any_cuda_work(lib):
    """Invalid descriptors are rejected with negative codes whether or not a GPU is present."""
    from cvar_b200 import _lib
    from cvar_b200.axis import build_axis
    x, dx = build_axis(64, "single")
    h = C.c_void_p()

    def create(mutate, xs=x):
        d = _lib.CvarDesc()
        lib.cvar_desc_default(C.byref(d))
        d.n, d.rho = 64, 0.5
        mutate(d)
        return lib.cvar_plan_create(C.byref(d), C.c_void_p(xs.ctypes.data), C.c_void_p(dx.ctypes.data), None, -1, C.byref(h))

    assert create(lambda d: setattr(d, "copula", 7)) == -2
    assert create(lambda d: setattr(d, "n", 1)) == -3
    assert create(lambda d: None, xs=np.ascontiguousarray(x[::-1])) == -3
    assert create(lambda d: setattr(d, "rho", 1.0)) == -4
    assert create(lambda d: setattr(d, "w0", 0.0)) == -4
    assert create(lambda d: setattr(d, "struct_size", 8)) == -7
    assert create(lambda d: (setattr(d, "marginal", 1), setattr(d, "q", 3))) == -1      # mixture without sigma states
    assert lib.cvar_plan_create(None, None, None, None, -1, C.byref(h)) == -1


def test_no_cpu_fallback_without_a_device(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from cvar_b200.backend import VarPlan
    from cvar_b200._lib import CvarError
    from cvar_b200.inputs import make_inputs
    inp = make_inputs("gaussian", "single", 64, sigma=np.ones((2, 2)))
    with pytest.raises(CvarError) as ei:
        VarPlan(inp)
    assert ei.value.status == -6


def test_product_package_never_imports_the_oracle():
    pkg = REPO / "copula-msm-and-copula-garch-var_b200"
    offenders = [str(p) for p in pkg.rglob("*.py") if re.search(r"^\s*(from|import)\s+oracle\b", p.read_text(), flags=re.M)]
    offenders += [str(p) for p in pkg.rglob("*.cu*") if "oracle" in p.read_text()]
    assert not offenders


def test_issued_fp64_per_cell_constants_match_the_sass():
    """bench.py's `roofline.cell_pipe_frac` multiplies the cell counter by workmodel.FP64_PER_CELL: re-count those
    constants in the SASS of the built library (tools/sass_loops.py lists every loop of a kernel with its FP64 count)."""
    import shutil
    import subprocess
    import sys
    from cvar_b200.build import LIB_PATH
    from cvar_b200.workmodel import CELLS_IN_FLIGHT, FP64_PER_CELL

    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not Path(cuobjdump).exists():
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", str(LIB_PATH)], capture_output=True, text=True, check=True).stdout
    for kv, (fast, slow) in FP64_PER_CELL.items():
        loops = subprocess.run([sys.executable, str(REPO / "tools" / "sass_loops.py"), f"solve_kernelILi{kv}ELb0", "1000"],
                               input=sass, capture_output=True, text=True, check=True).stdout
        fp64_counts = {int(m.group(1)) for m in re.finditer(r"FP64\s+(\d+),", loops)}
        for per_cell in {fast, slow}:
            assert per_cell * CELLS_IN_FLIGHT[kv] in fp64_counts, (kv, per_cell, sorted(fp64_counts)[-12:])


def test_dimension_decision_is_a_status_code(lib):
    """f4: portfolios with dim != 2 are refused, explicitly (cvar_check_dim / CVAR_ERR_DIM), in the C ABI and above it."""
    from cvar_b200 import _lib
    assert lib.cvar_check_dim(2) == 0
    for dim in (1, 3, 5):
        assert lib.cvar_check_dim(dim) == -10
    assert b"two-asset" in lib.cvar_strerror(-10)
    with pytest.raises(NotImplementedError, match="two-asset"):
        _lib.check_dim(3)
