"""Plugin contract of the VaR calculators (mirror of the reference's utils/calc_var_ABC.py:25-111).

A calculator is what `ValueAtRiskCalculationFactory.create_var_calculator` returns and what
`ValueAtRiskCalcualtion` consumes.  The four abstract methods are the reference's; the remaining hooks the
driver reads (`copula_integrations_params`, `unpack_copula_params`, `copula_density`, `integrated_function`)
are duck-typed there (utils/calc_var_class.py:43-45,62,68,78,81,86) and documented here.
"""
from abc import ABC, abstractmethod


class _SharedCache:
    """Process-global memo dictionaries keyed by ticker/date (reference: calc_var_ABC.py:4-22)."""
    cache = {}


class SharedCacheCopulaMSMVaR(_SharedCache):
    cache = {}


class SharedCacheCopulaGarchVaR(_SharedCache):
    cache = {}


class SharedCacheCopulaMRVaR(_SharedCache):
    cache = {}


class VaRCalculationMethod(ABC):
    """Interface of a (copula, marginal model) VaR calculator.

    model_params_insample(in_sample_dict, *a, **kw)            -> per-ticker fitted parameters
    calculate_marginals_and_densities_in_sample(in_sample_dict, in_sample_params, *a, **kw)
                                                               -> (marginals[N,dim], densities[N,dim], vol_states|None)
    copula_or_correl_params_insample(marginals, densities)     -> dict of fitted copula parameters
    integration_params_retrieval(dim, rolling_windows_dict, in_sample_params, num_points, vol_states)
                                                               -> (integrations_params_t, integrations_params_static,
                                                                   grids_generations_params)
    """

    @abstractmethod
    def model_params_insample(self):
        ...

    @abstractmethod
    def calculate_marginals_and_densities_in_sample(self):
        ...

    @abstractmethod
    def copula_or_correl_params_insample(self):
        ...

    @abstractmethod
    def integration_params_retrieval(self):
        ...


class OutOfScopeStage(NotImplementedError):
    """Raised by the in-sample fitting stages, which this backend does not re-implement.

    The B200 backend covers the per-day VaR solve.  Fit the marginal models and the copula with the
    reference's own (CPU, one-off) code and hand the resulting forecasts to
    `ValueAtRiskCalcualtion.from_forecasts(...)`, or patch the reference in place with
    `cvar_b200.dropin.install()`.
    """
