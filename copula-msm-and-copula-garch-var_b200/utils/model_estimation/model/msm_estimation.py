"""Binomial MSM marginal adapter (mirror of utils/model_estimation/model/msm_estimation.py:10-462)."""
import numpy as np

from cvar_b200 import msm_layout
from utils.calc_var_ABC import OutOfScopeStage, VaRCalculationMethod


class MSMEstimation(VaRCalculationMethod):
    marginal_family = "mixture"

    def __init__(self, state_prob_forecasts=None):
        """state_prob_forecasts: optional (dim, T, 2**k) filtered state probabilities produced elsewhere."""
        self.state_prob_forecasts = None if state_prob_forecasts is None else np.asarray(state_prob_forecasts, float)

    # ---- in-sample stages: not part of the GPU hot path ------------------------------------------
    def model_params_insample(self, in_sample_dict, k=None, *args, **kwargs):
        raise OutOfScopeStage("MSM parameter fitting (basin hopping) is outside the GPU hot path")

    def calculate_marginals_and_densities_in_sample(self, in_sample_dict, in_sample_params, k=None, *args, **kwargs):
        raise OutOfScopeStage("MSM in-sample PIT values are outside the GPU hot path")

    def copula_or_correl_params_insample(self):
        pass

    def density_function(self):
        pass

    def forecasts_array(self, rolling_windows_dict=None, in_sample_params=None, k=None):
        if self.state_prob_forecasts is None:
            raise OutOfScopeStage("MSM rolling-window Hamilton filtering is outside the GPU hot path; pass "
                                  "state_prob_forecasts= to the adapter or use ValueAtRiskCalcualtion.from_forecasts")
        return self.state_prob_forecasts

    # ---- hot-path input layout ---------------------------------------------------------------------
    def integration_params_retrieval(self, dim, rolling_windows_dict, in_sample_params, num_points, vol_state_array):
        forecasts_array = self.forecasts_array(rolling_windows_dict, in_sample_params)
        forecasts_by_states, unique_vol_states = self.sum_forecast_by_state(vol_state_array, forecasts_array)
        densities, x, dx = self.compute_normal_densities(unique_vol_states, num_points)
        integrations_params_t = forecasts_by_states, self.compute_forecast_combinations(forecasts_by_states)
        grids_generations_params = densities, x, dx, self.create_vol_combinations(unique_vol_states)
        return integrations_params_t, unique_vol_states, grids_generations_params

    @staticmethod
    def sum_forecast_by_state(vol_state_array, forecasts_array, tol=1e-6):
        return msm_layout.merge_states(vol_state_array, forecasts_array, tol)

    @staticmethod
    def compute_normal_densities(unique_vol_states_array, num_points, x_min=-5, x_max=5):
        return msm_layout.state_densities(unique_vol_states_array, num_points)

    @staticmethod
    def create_vol_combinations(unique_vol_states):
        dim, q = np.asarray(unique_vol_states).shape
        return msm_layout.state_index_pairs(dim, q)

    @staticmethod
    def compute_forecast_combinations(summed_forecasts):
        return msm_layout.pair_probabilities(summed_forecasts)

    @staticmethod
    def integrated_function(grids, step_sizes, copula_params, integrations_params_i, integrations_params_static,
                            copula_density, unpack_copula_params):
        raise OutOfScopeStage("the B200 backend never materialises per-state-pair weight columns; "
                              "use ValueAtRiskCalcualtion.compute_integral")
