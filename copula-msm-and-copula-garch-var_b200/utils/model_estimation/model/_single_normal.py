"""Common part of the two single-normal marginal adapters (GARCH and Kalman mean-reverting): on the hot
path they are indistinguishable -- one forecast sigma per asset per day."""
import numpy as np

from cvar_b200.axis import build_axis
from utils.calc_var_ABC import OutOfScopeStage, VaRCalculationMethod


class SingleNormalEstimation(VaRCalculationMethod):
    marginal_family = "single"
    model_name = "single-normal"

    def __init__(self, sigma_forecasts=None):
        """sigma_forecasts: optional (T, dim) array of one-step vol forecasts produced elsewhere."""
        self.sigma_forecasts = None if sigma_forecasts is None else np.asarray(sigma_forecasts, dtype=float)

    # ---- in-sample stages: not part of the GPU hot path ------------------------------------------
    def model_params_insample(self, in_sample_dict, *args, **kwargs):
        raise OutOfScopeStage(f"{self.model_name} parameter fitting is outside the GPU hot path")

    def calculate_marginals_and_densities_in_sample(self, in_sample_dict, in_sample_params, *args, **kwargs):
        raise OutOfScopeStage(f"{self.model_name} in-sample PIT values are outside the GPU hot path")

    def copula_or_correl_params_insample(self):
        pass

    def forecasts_array(self):
        pass

    def sum_forecast_by_state(self):
        pass

    def density_function(self):
        pass

    def compute_forecast(self, rolling_windows_dict, in_sample_params):
        if self.sigma_forecasts is None:
            raise OutOfScopeStage(f"{self.model_name} rolling-window forecasts are outside the GPU hot path; "
                                  "pass sigma_forecasts= to the adapter or use ValueAtRiskCalcualtion.from_forecasts")
        return [self.sigma_forecasts]

    # ---- hot-path input layout ---------------------------------------------------------------------
    @staticmethod
    def compute_normal_densities(dim, num_points, x_min=-5, x_max=5):
        """(densities = ones(dim, 1, n), x, dx): the single-normal axis (outer n//8, middle n//5 points)."""
        x, dx = build_axis(num_points, "single", x_min, x_max)
        return np.ones((dim, 1, num_points)), x, dx

    def integration_params_retrieval(self, dim, rolling_windows_dict, in_sample_params, num_points, vol_state_array):
        densities, x, dx = self.compute_normal_densities(dim, num_points)
        grids_generations_params = densities, x, dx, np.zeros((1, dim))
        return self.compute_forecast(rolling_windows_dict, in_sample_params), None, grids_generations_params

    @staticmethod
    def integrated_function(grids, step_sizes, copula_params, integrations_params_i, integrations_params_static,
                            copula_density, unpack_copula_params):
        """Integrand on an explicit point list (reference: integration_functions/garch_integration_function.py).
        Evaluated on the GPU for API compatibility; `calc_var` itself never materialises point lists."""
        from cvar_b200.density import integrand_single_gpu
        nu, corr = unpack_copula_params(copula_params)
        return integrand_single_gpu(grids, step_sizes, integrations_params_i[0] if isinstance(
            integrations_params_i, (list, tuple)) else integrations_params_i, nu, corr)
