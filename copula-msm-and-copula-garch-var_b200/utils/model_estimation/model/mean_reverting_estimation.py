"""Kalman mean-reverting log-vol marginal adapter
(mirror of utils/model_estimation/model/mean_reverting_estimation.py:11-252)."""
import numpy as np

from utils.calc_var_ABC import OutOfScopeStage
from utils.model_estimation.model._single_normal import SingleNormalEstimation


class MeanRevertingEstimation(SingleNormalEstimation):
    model_name = "Kalman mean-reverting"

    def compute_forecast(self, rolling_windows_dict, in_sample_params):
        """[sigma[T, dim]] = exp(last predicted log-vol) of the unscented filter per rolling window, on the GPU
        (cvar_b200.forecast.kalman_forecast; reference: mean_reverting_estimation.py:192-232 ->
        kalman_mean_reverting/forecast.py:5-12).  Parameters: in_sample_params[ticker]['optimal_params'] = {a, l, q}."""
        if self.sigma_forecasts is not None:
            return [self.sigma_forecasts]
        if not rolling_windows_dict or not in_sample_params:
            raise OutOfScopeStage("Kalman forecasts need rolling windows and fitted parameters (the EM fit itself is outside "
                                  "the GPU hot path); or pass sigma_forecasts= to the adapter")
        from cvar_b200.forecast import rolling_series
        tickers = list(in_sample_params)
        windows = [np.array([w[t] for w in rolling_windows_dict.values()], dtype=float) for t in tickers]
        N = windows[0].shape[1]
        series = [rolling_series(w) for w in windows]
        if all(s is not None for s in series):
            return self.forecast_from_series(np.array(series), in_sample_params, N)
        return self.forecast_from_series(np.array([w.reshape(-1) for w in windows]), in_sample_params, N, window_stride=N)

    @staticmethod
    def forecast_from_series(series, in_sample_params, N, window_stride=1, **_):
        """[sigma[T, dim]] from centred return series (dim, (T-1)*window_stride + N)."""
        from cvar_b200.forecast import kalman_forecast
        a, l, q = ([in_sample_params[t]["optimal_params"][key] for t in in_sample_params] for key in ("a", "l", "q"))
        sigma, info = kalman_forecast(series, a, l, q, N, window_stride=window_stride)
        if info["failed"]:
            raise FloatingPointError("the unscented filter collapsed (normalising constant <= 1e-10) in at least one window")
        return [sigma]
