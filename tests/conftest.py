"""Shared test plumbing: import paths, the `gpu` marker, golden-vector loading."""
import sys
from pathlib import Path

import numpy as np
import pytest

REPO = Path(__file__).resolve().parent.parent
PKG_ROOT = REPO / "copula-msm-and-copula-garch-var_b200"
for p in (str(PKG_ROOT), str(REPO), str(REPO / "baseline")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN_DIR = REPO / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # the CUDA library is a build artefact (git-ignored): compile it if this checkout does not have it yet
    from cvar_b200.build import LIB_PATH, build_library
    if not LIB_PATH.exists():
        try:
            build_library()
        except Exception as exc:      # no nvcc on this checkout: the pure-CPU tests (oracle, goldens, gloo) still run;
            import warnings           # ABI and gpu-marked tests fail or skip on their own when they load the library
            warnings.warn(f"libcvar_b200.so is missing and could not be built: {exc}")


NON_SOLVE_GOLDENS = {"forecast_producers", "ambiguous_exit_seeds"}


def golden_names():
    """Golden cases of the VaR solve (one .npz per case); other fixtures live in NON_SOLVE_GOLDENS."""
    return sorted(p.stem for p in GOLDEN_DIR.glob("*.npz") if p.stem not in NON_SOLVE_GOLDENS)


def load_golden(name):
    """(HotPathInputs, npz dict) of one golden case produced by tests/golden/make_golden.py."""
    from cvar_b200.inputs import HotPathInputs

    g = dict(np.load(GOLDEN_DIR / f"{name}.npz"))
    marginal = str(g["marginal"])
    kw = dict(sigma=g["sigma"]) if marginal == "single" else dict(probs=g["probs"], sigma_states=g["sigma_states"])
    inp = HotPathInputs(copula=str(g["copula"]), marginal=marginal, n=int(g["n"]), x=g["ref_x"], dx=g["ref_dx"],
                        weights=g["weights"], rho=float(g["rho"]), nu=float(g["nu"]), theta=float(g["theta"]),
                        ptf_mean=float(g["ptf_mean"]), **kw)
    return inp, g


@pytest.fixture(scope="session")
def cuda_device():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")
