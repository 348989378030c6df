"""`ValueAtRiskCalcualtion` on the B200 backend (mirror of the reference's utils/calc_var_class.py:8-309).

Same constructor signature, attribute names and public methods as the reference (spelling included).
`calc_var` hands the whole bracket + bisection to the CUDA library (one launch for all days) instead of
driving ~23 host-side `compute_integral` round trips; `compute_integral`, `adjust_integral` and
`bisection_algorithm` stay available with the reference's semantics, built on GPU strip masses.
"""
import time

import numpy as np

from cvar_b200 import _lib
from cvar_b200.backend import VarPlan
from cvar_b200.inputs import HotPathInputs
from utils.calc_var_ABC import OutOfScopeStage


def hot_path_inputs_from_attributes(obj, copula_family=None, marginal_family=None) -> HotPathInputs:
    """Collect the solve's inputs from the reference-layout attributes of a calculator driver object
    (`copula_params`, `integrations_params_t`, `integrations_params_static`, `grids_generations_params`,
    `weights`, `ptf_mean`, `num_points`): works for this module's class and for the reference's own."""
    method = obj.VaRCalculationMethod
    copula = copula_family or getattr(method, "copula_family", None) or {
        "GaussianCopulaVaR": "gaussian", "StudentCopulaVaR": "student", "PlackettCopulaVaR": "plackett",
    }[type(method).__name__]
    marginal = marginal_family or ("single" if obj.integrations_params_static is None else "mixture")
    _, x, dx, _ = obj.grids_generations_params
    kw = dict(copula=copula, marginal=marginal, n=int(obj.num_points), x=x, dx=dx,
              weights=np.asarray(obj.weights, float), ptf_mean=float(obj.ptf_mean))
    cp = obj.copula_params
    if copula == "gaussian":
        kw["rho"] = float(np.atleast_1d(cp)[0])
    elif copula == "student":
        kw["nu"], kw["rho"] = float(cp[0]), float(cp[1])
    else:
        kw["theta"] = float(cp)
    if marginal == "single":
        kw["sigma"] = np.asarray(obj.integrations_params_t[0], float)
    else:
        kw["probs"] = np.asarray(obj.integrations_params_t[0], float)
        kw["sigma_states"] = np.asarray(obj.integrations_params_static, float)
    _lib.check_dim(int(getattr(obj, "dim", 2)))
    _lib.check_dim(len(kw["weights"]))
    return HotPathInputs(**kw)


class ValueAtRiskCalcualtion:
    def __init__(self, tickers, start_date, in_sample_data_num, VaRCalculationMethod, end_date=None, num_points=100,
                 weights=np.array([0.5, 0.5]), *args, **kwargs):
        self.num_points = num_points
        self.tickers = tickers
        self.start_date = start_date
        self.in_sample_data_num = in_sample_data_num
        self.VaRCalculationMethod = VaRCalculationMethod
        self.end_date = end_date
        self.weights = weights

        (self.in_sample_dict, self.rolling_windows_dict, self.mean_returns, self.end_date, self.out_sample_data,
         self.out_sample_N, self.dim, self.ptf_mean) = self.get_in_sample_data()

        self.in_sample_params = self.retrieve_param_in_sample(*args, **kwargs)
        self.marginals, self.densities, self.vol_states_array = self.calc_marg_and_densities(*args, **kwargs)
        self.copula_params = self.calc_copula_params()
        self.integrations_params_t, self.integrations_params_static, self.grids_generations_params = (
            self.integration_params_retrieval())
        self._bind_hooks()

    # ---- alternate construction from ready-made forecasts (no data download, no fitting) -------------
    @classmethod
    def from_forecasts(cls, VaRCalculationMethod, copula_params, *, sigma=None, state_probs=None, vol_states=None,
                       probs_by_state=None, sigma_states=None, num_points=100, weights=np.array([0.5, 0.5]),
                       ptf_mean=0.0, out_sample_data=None):
        """Build the driver directly from per-day forecast parameters.

        single-normal models (GARCH, Kalman):  sigma[T, 2]
        MSM: either raw `state_probs[2, T, 2**k]` + `vol_states[2, 2**k]` (merged here exactly like the
        reference's adapter) or already merged `probs_by_state[T, 2, q]` + `sigma_states[2, q]`.
        """
        self = object.__new__(cls)
        self.num_points = int(num_points)
        self.tickers = self.start_date = self.end_date = self.in_sample_data_num = None
        self.VaRCalculationMethod = VaRCalculationMethod
        self.weights = np.asarray(weights, float)
        self.dim = 2
        self.ptf_mean = float(ptf_mean)
        self.in_sample_dict = self.rolling_windows_dict = self.mean_returns = None
        self.in_sample_params = self.marginals = self.densities = None
        self.out_sample_data = out_sample_data
        self.copula_params = copula_params
        m = VaRCalculationMethod
        if m.marginal_family == "single":
            if sigma is None:
                raise ValueError("single-normal marginals need sigma[T, 2]")
            sigma = np.ascontiguousarray(sigma, dtype=float)
            densities, x, dx = m.compute_normal_densities(self.dim, self.num_points)
            self.vol_states_array = None
            self.integrations_params_t = [sigma]
            self.integrations_params_static = None
            self.grids_generations_params = (densities, x, dx, np.zeros((1, self.dim)))
            self.out_sample_N = sigma.shape[0]
        else:
            if probs_by_state is None:
                if state_probs is None or vol_states is None:
                    raise ValueError("MSM marginals need state_probs + vol_states or probs_by_state + sigma_states")
                probs_by_state, sigma_states = m.sum_forecast_by_state(np.asarray(vol_states), np.asarray(state_probs))
            probs_by_state = np.ascontiguousarray(probs_by_state, dtype=float)
            sigma_states = np.ascontiguousarray(sigma_states, dtype=float)
            densities, x, dx = m.compute_normal_densities(sigma_states, self.num_points)
            self.vol_states_array = vol_states
            self.integrations_params_t = (probs_by_state, m.compute_forecast_combinations(probs_by_state))
            self.integrations_params_static = sigma_states
            self.grids_generations_params = (densities, x, dx, m.create_vol_combinations(sigma_states))
            self.out_sample_N = probs_by_state.shape[0]
        self._bind_hooks()
        return self

    @classmethod
    def from_returns(cls, VaRCalculationMethod, returns, in_sample_data_num, in_sample_params, copula_params, *,
                     num_points=100, weights=np.array([0.5, 0.5]), k=None):
        """Returns -> forecasts -> ready-to-solve driver, with the forecast stage on the GPU.

        returns            : DataFrame (columns = tickers) or array (days, 2) of percent log-returns
        in_sample_data_num : N, length of the in-sample period and of every rolling window
        in_sample_params   : fitted model parameters in the reference's own layout, one entry per ticker/column
                             ({name: {'optimal_params': {...}}}: GARCH best_pq/best_params, MSM m_0/sig/b/gamma,
                             Kalman a/l/q) -- fitting itself is outside this backend
        copula_params      : as `copula_integrations_params` returns them

        Follows the reference's data preparation (data_loader/load_data.py:105-137): centring by the in-sample mean,
        rolling window i = centred returns[i : i+N], `ptf_mean = sum(mean_returns * weights)`,
        `out_sample_data = returns[N:]`.
        """
        N = int(in_sample_data_num)
        values = returns.to_numpy(dtype=float) if hasattr(returns, "to_numpy") else np.asarray(returns, dtype=float)
        if values.ndim != 2 or values.shape[1] != 2:
            raise NotImplementedError("the B200 backend covers two-asset portfolios")
        if values.shape[0] <= N:
            raise ValueError(f"Not enough returns for in-sample estimation. Required: {N + 1}, Available: {values.shape[0]}")
        weights = np.asarray(weights, float)
        mean_returns = values[:N].mean(axis=0)
        T = values.shape[0] - N
        series = np.ascontiguousarray((values - mean_returns)[: T + N - 1].T)
        m = VaRCalculationMethod
        kw = dict(num_points=num_points, weights=weights, ptf_mean=float(np.sum(mean_returns * weights)),
                  out_sample_data=returns.iloc[N:] if hasattr(returns, "iloc") else values[N:])
        if m.marginal_family == "single":
            (sigma,) = m.estimation_method.forecast_from_series(series, in_sample_params, N)
            self = cls.from_forecasts(m, copula_params, sigma=sigma, **kw)
        else:
            pbs, sig = m.estimation_method.forecast_from_series(series, in_sample_params, N, k=k)
            self = cls.from_forecasts(m, copula_params, probs_by_state=pbs, sigma_states=sig, **kw)
        self.in_sample_data_num, self.in_sample_params = N, in_sample_params
        self.mean_returns = dict(zip(getattr(returns, "columns", range(2)), mean_returns))
        return self

    def _bind_hooks(self):
        m = self.VaRCalculationMethod
        self.copula_function = m.copula_density
        self.unpack_copula_params = m.unpack_copula_params
        self.integrated_function = m.integrated_function

    # ---- constructor stages (same call sequence as the reference, :47-93) -----------------------------
    def get_in_sample_data(self):
        try:
            from data_loader.load_data import IndexReturnsRetriever   # the reference's loader, if on sys.path
        except ImportError as exc:
            raise OutOfScopeStage(
                "market-data download (yfinance) is outside the GPU hot path and is not shipped; put the reference's "
                "`data_loader` on sys.path or use ValueAtRiskCalcualtion.from_forecasts(...)") from exc
        retriever = IndexReturnsRetriever(tickers=self.tickers, start_date=self.start_date, N=self.in_sample_data_num,
                                          weights=self.weights, end_date=self.end_date)
        return retriever.get_insample_data()

    def retrieve_param_in_sample(self, *args, **kwargs):
        return self.VaRCalculationMethod.model_params_insample(self.in_sample_dict, *args, **kwargs)

    def calc_marg_and_densities(self, *args, **kwargs):
        return self.VaRCalculationMethod.calculate_marginals_and_densities_in_sample(
            self.in_sample_dict, self.in_sample_params, *args, **kwargs)

    def calc_copula_params(self):
        best_fit = self.VaRCalculationMethod.copula_or_correl_params_insample(self.marginals, self.densities)
        return self.VaRCalculationMethod.copula_integrations_params(best_fit)

    def integration_params_retrieval(self):
        return self.VaRCalculationMethod.integration_params_retrieval(
            self.dim, self.rolling_windows_dict, self.in_sample_params, self.num_points, self.vol_states_array)

    # ---- GPU plumbing -----------------------------------------------------------------------------
    def hot_path_inputs(self) -> HotPathInputs:
        return hot_path_inputs_from_attributes(self)

    def _plan(self, first_guess=-3, second_guess=(-3.5, -2), inputs=None) -> VarPlan:
        """Plan for the object's CURRENT run constants (copula parameters, weights, grid, vol levels, guesses): like
        the reference, which re-reads its attributes on every calc_var, a refit or a changed `num_points` takes effect
        on the next call; an unchanged object reuses its plan."""
        from cvar_b200.dropin import _plan_for

        return _plan_for(self, inputs if inputs is not None else self.hot_path_inputs(), first_guess, second_guess)

    # ---- the hot path ------------------------------------------------------------------------------
    def calc_var(self, obj_var=0.05, first_guess=-3, second_guess=(-3.5, -2)):
        """VaR level per out-of-sample day at tail probability `obj_var` (np.ndarray[T] = solved q + ptf_mean)."""
        return self.calc_var_multi([obj_var], first_guess, second_guess)[0]

    def calc_var_multi(self, obj_vars, first_guess=-3, second_guess=(-3.5, -2)):
        """Several tail probabilities in ONE launch (they share the per-day axis work): array (len(obj_vars), T).
        Each row equals what `calc_var(obj_var)` returns on its own."""
        start = time.time()
        inp = self.hot_path_inputs()
        res = self._plan(first_guess, second_guess, inp).solve(inp.day_params(), np.asarray(obj_vars, float),
                                                          ptf_mean=self.ptf_mean)
        self.last_solve = res
        self.last_calc_var_seconds = time.time() - start
        return res.var

    def compute_integral(self, bounds):
        """Strip mass per day for `bounds[T, 2]` (the reference's grid build + joblib fan-out, :179-212)."""
        inp = self.hot_path_inputs()
        return self._plan(inputs=inp).strip_mass(inp.day_params(), np.asarray(bounds, float))

    @staticmethod
    def adjust_integral(new_result, prev_results, bounds, prev_upper):
        """prev + new where the strip starts at prev_upper, else prev - new (exact float ==, :214-248)."""
        bounds = np.asarray(bounds, float)
        add = bounds[:, 0] == np.asarray(prev_upper, float)
        return np.where(add, prev_results + new_result, prev_results - new_result)

    def bisection_algorithm(self, obj_var, bisection_bounds, prev_result, upper_stack, prev_upper, tolerance=1e-6):
        """Host-driven vectorised bisection with the reference's semantics (:250-309), one GPU strip-mass
        launch per iteration.  `calc_var` does NOT use this (its loop runs in-kernel); it is kept for API
        parity and as an independent cross-check of the in-kernel state machine."""
        lower = np.array(bisection_bounds[:, 0], float)
        upper = np.array(bisection_bounds[:, 1], float)
        upper_stack = np.asarray(upper_stack, bool)
        prev_result = np.asarray(prev_result, float)
        prev_upper = np.asarray(prev_upper, float)
        while np.any(upper - lower > tolerance):
            mid = (lower + upper) / 2
            bounds = np.where(upper_stack[:, None], np.column_stack((lower, mid)), np.column_stack((mid, upper)))
            current = self.adjust_integral(self.compute_integral(bounds), prev_result, bounds, prev_upper)
            if np.all(current == 0):
                break
            upper_stack = current < obj_var
            lower = np.where(upper_stack, mid, lower)
            upper = np.where(upper_stack, upper, mid)
            prev_result, prev_upper = current, mid
        return (lower + upper) / 2
