"""GARCH(p,q) marginal adapter (mirror of utils/model_estimation/model/garch_estimation.py:11-251)."""
import numpy as np

from utils.calc_var_ABC import OutOfScopeStage
from utils.model_estimation.model._single_normal import SingleNormalEstimation


class GarchEstimation(SingleNormalEstimation):
    model_name = "GARCH"

    def compute_forecast(self, rolling_windows_dict, in_sample_params):
        """[sigma[T, dim]]: one-step volatility forecast at the end of every rolling window, on the GPU
        (cvar_b200.forecast.garch_forecast) instead of the reference's per-date Python loop
        (garch_estimation.py:190-231 -> garch/forecast.py:5-18).  Parameters are read from
        `in_sample_params[ticker]['optimal_params']` = {'best_pq': (p, q), 'best_params': [omega, alpha.., beta..]}."""
        if self.sigma_forecasts is not None:
            return [self.sigma_forecasts]
        if not rolling_windows_dict or not in_sample_params:
            raise OutOfScopeStage("GARCH forecasts need rolling windows and fitted parameters (fitting itself is outside "
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
        """[sigma[T, dim]] from centred return series (dim, (T-1)*window_stride + N); one row of `series` per ticker of
        `in_sample_params` (in its order)."""
        from cvar_b200.forecast import garch_forecast
        omega, alphas, betas = [], [], []
        for t in in_sample_params:
            prm = in_sample_params[t]["optimal_params"]
            p = prm["best_pq"][0]
            best = np.asarray(prm["best_params"], dtype=float)
            omega.append(best[0]); alphas.append(best[1:p + 1]); betas.append(best[p + 1:])
        sigma, _ = garch_forecast(series, omega, alphas, betas, N, window_stride=window_stride)
        return [sigma]
