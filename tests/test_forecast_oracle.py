"""Forecast-producer oracle against the golden vectors of the unmodified reference (CPU)."""
import numpy as np
import pytest

from conftest import GOLDEN_DIR
from oracle import forecast_oracle as fo

G = dict(np.load(GOLDEN_DIR / "forecast_producers.npz"))
MSM = sorted({k.split("__")[0] for k in G if k.startswith("msm")})
GARCH = sorted({k.split("__")[0] for k in G if k.startswith("garch")})
KALMAN = sorted({k.split("__")[0] for k in G if k.startswith("kalman")})


def case(name):
    return {k.split("__")[1]: v for k, v in G.items() if k.startswith(name + "__")}


@pytest.mark.parametrize("name", MSM)
def test_msm_filter_oracle_matches_reference(name):
    c = case(name)
    got = fo.msm_forecast(c["series"], int(c["k"]), float(c["m0"]), float(c["sigma_bar"]), float(c["b"]), float(c["gamma"]),
                          int(c["N"]), int(c["T"]))
    assert got.shape == c["ref"].shape
    np.testing.assert_allclose(got, c["ref"], rtol=1e-13, atol=1e-300)
    np.testing.assert_allclose(got.sum(axis=1), 1.0, atol=1e-13)


@pytest.mark.parametrize("name", GARCH)
def test_garch_forecast_oracle_matches_reference(name):
    c = case(name)
    got = fo.garch_forecast(c["series"], float(c["omega"]), c["alpha"], c["beta"], int(c["N"]), int(c["T"]))
    assert got.tobytes() == c["ref"].tobytes()


@pytest.mark.parametrize("name", KALMAN)
def test_kalman_forecast_oracle_matches_reference(name):
    c = case(name)
    got = fo.kalman_forecast(c["series"], float(c["a"]), float(c["l"]), float(c["q"]), int(c["N"]), int(c["T"]))
    np.testing.assert_allclose(got, c["ref"], rtol=1e-13)


def test_kronecker_structure_of_the_transition_matrix():
    """The dense matrix of the reference is the Kronecker product of its k 2x2 factors (component 0 slowest)."""
    from cvar_b200.msm_layout import msm_stay_probs, msm_transition_matrix
    k, m0, b, gamma = 5, 0.45, 2.5, 0.2
    p = msm_stay_probs(k, b, gamma)
    kron = np.array([[1.0]])
    for c in range(k):
        kron = np.kron(kron, np.array([[p[c], 1 - p[c]], [1 - p[c], p[c]]]))
    np.testing.assert_allclose(kron, msm_transition_matrix(k, m0, b, gamma), rtol=1e-14)
    np.testing.assert_allclose(kron, fo.msm_tables(k, m0, 1.0, b, gamma)[1], rtol=1e-14)
