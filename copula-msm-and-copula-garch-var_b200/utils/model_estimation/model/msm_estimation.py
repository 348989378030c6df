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
        """(dim, T, 2**k) filtered state probabilities at the end of every rolling window.

        Runs on the GPU (cvar_b200.forecast.msm_forecast: one warp per window, Kronecker-factored transition)
        instead of the reference's Python loop over dates around a dense numba filter
        (msm_estimation.py:143-203).  Model parameters come from `in_sample_params[ticker]['optimal_params']`
        = {'m_0', 'sig', 'b', 'gamma'} exactly as the reference stores them."""
        if self.state_prob_forecasts is not None:
            return self.state_prob_forecasts
        if not rolling_windows_dict or not in_sample_params or k is None:
            raise OutOfScopeStage("MSM forecasts need rolling windows, fitted parameters and k (fitting itself is outside "
                                  "the GPU hot path); or pass state_prob_forecasts= to the adapter")
        from cvar_b200.forecast import rolling_series
        tickers = list(in_sample_params)
        windows = [np.array([w[t] for w in rolling_windows_dict.values()], dtype=float) for t in tickers]
        N = windows[0].shape[1]
        series = [rolling_series(w) for w in windows]
        if all(s is not None for s in series):
            return self.state_probs_from_series(np.array(series), in_sample_params, N, int(k))
        # windows that do not overlap like a rolling series: filter each one on its own
        return self.state_probs_from_series(np.array([w.reshape(-1) for w in windows]), in_sample_params, N, int(k), window_stride=N)

    @staticmethod
    def _msm_params(in_sample_params):
        from cvar_b200.forecast import MsmParams
        return [MsmParams(m0=v["optimal_params"]["m_0"], sigma_bar=v["optimal_params"]["sig"], b=v["optimal_params"]["b"],
                          gamma=v["optimal_params"]["gamma"]) for v in in_sample_params.values()]

    @classmethod
    def state_probs_from_series(cls, series, in_sample_params, N, k, window_stride=1):
        """(dim, T, 2**k) filtered state probabilities from centred return series (dim, (T-1)*window_stride + N)."""
        from cvar_b200.forecast import msm_forecast
        _, _, state_probs, info = msm_forecast(series, cls._msm_params(in_sample_params), k, N, window_stride=window_stride,
                                               return_state_probs=True)
        if info["degenerate"]:
            raise FloatingPointError("MSM filter degenerated (zero normalising constant) in at least one window")
        return state_probs

    @classmethod
    def forecast_from_series(cls, series, in_sample_params, N, window_stride=1, k=None, **_):
        """(probs_by_state[T, dim, q], sigma_states[dim, q]) straight from centred return series -- the merged layout
        the solve consumes, produced by the filter kernel itself."""
        from cvar_b200.forecast import msm_forecast
        if k is None:
            raise ValueError("MSM forecasts need k (number of multiplier components)")
        pbs, sig, info = msm_forecast(series, cls._msm_params(in_sample_params), int(k), N, window_stride=window_stride)
        if info["degenerate"]:
            raise FloatingPointError("MSM filter degenerated (zero normalising constant) in at least one window")
        return pbs, sig

    # ---- hot-path input layout ---------------------------------------------------------------------
    def integration_params_retrieval(self, dim, rolling_windows_dict, in_sample_params, num_points, vol_state_array):
        # the reference derives k as int(sqrt(2**k)) (msm_estimation.py:125, quirk Q12), which is only right for
        # k in {1, 2, 4}; the state count is 2**k, so k is its base-2 logarithm
        k = int(round(np.log2(np.asarray(vol_state_array).shape[1])))
        forecasts_array = self.forecasts_array(rolling_windows_dict, in_sample_params, k)
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
