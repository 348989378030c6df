"""Shared body of the three copula calculators: delegation of the model-side methods to the wrapped
marginal-model adapter (the reference repeats this block in every copula class,
utils/model_estimation/copula/*_estimation.py)."""
import numpy as np

from utils.calc_var_ABC import OutOfScopeStage, VaRCalculationMethod

_DELEGATED = (
    "calculate_marginals_and_densities_in_sample", "density_function", "forecasts_array", "model_params_insample",
    "sum_forecast_by_state", "compute_normal_densities", "create_vol_combinations", "compute_forecast_combinations",
    "integrated_function", "integration_params_retrieval",
)


def corr_from_rho(rho):
    """Symmetric correlation matrix from its strict upper triangle (row-major), as the reference unpacks it."""
    rho = np.atleast_1d(np.asarray(rho, dtype=float))
    dim = int((1 + np.sqrt(1 + 8 * len(rho))) / 2)
    corr = np.eye(dim)
    corr[np.triu_indices(dim, k=1)] = rho
    corr[np.tril_indices(dim, k=-1)] = rho
    return corr


class CopulaVaRBase(VaRCalculationMethod):
    """Composite calculator = copula family + marginal-model adapter (`estimation_method`)."""

    copula_family = None      # 'gaussian' | 'student' | 'plackett' : selects the CUDA cell kernel

    def __init__(self, estimation_method):
        self.estimation_method = estimation_method

    @property
    def marginal_family(self):
        return self.estimation_method.marginal_family

    @staticmethod
    def copula_or_correl_params_insample(marginals, densities):
        raise OutOfScopeStage("copula parameter fitting (IFM / L-BFGS-B) is outside the GPU hot path")


def _delegate(name):
    def method(self, *args, **kwargs):
        return getattr(self.estimation_method, name)(*args, **kwargs)
    method.__name__ = name
    return method


for _name in _DELEGATED:
    setattr(CopulaVaRBase, _name, _delegate(_name))
CopulaVaRBase.__abstractmethods__ = frozenset()
