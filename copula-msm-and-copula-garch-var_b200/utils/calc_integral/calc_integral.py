"""`calc_grids_and_integrals_results` on the B200 backend -- the lowest pure-function seam of the hot path
(mirror of the reference's utils/calc_integral/calc_integral.py:8-119).

Same sixteen arguments as the reference.  The reference materialises one nested grid per unique (lower, upper) pair
(create_grids.py:6-171) and fans the per-day integrand out over joblib (calc_integral.py:171-225); here the strip
masses of all days come from ONE launch of the strip-mass kernel (`cvar_strip_mass_host`), which never materialises
a grid.  Arguments that only exist to build those grids are checked for the values the kernel implements and are
otherwise unused:

* `var_function` must be the reference's portfolio-return function (integration_algo.py:20): membership is
  `x[j] <= (var - x[i] * w[1]) / w[0]`, evaluated with individually rounded operations;
* `upper_bound` is not read by the reference's membership either (create_grids.py:102-110); `lower_bound` is the clip
  of the inner axis (-5 in calc_var_class.py:201) and is honoured;
* `integrated_function`, `copula_density` and `unpack_copula_params` identify the marginal and copula families (the
  calculators' hooks carry `copula_family`; reference hooks are recognised by their qualified names).

There is no CPU fallback: without the CUDA library this raises.
"""
from collections import OrderedDict

import numpy as np

from cvar_b200 import _lib
from cvar_b200.backend import VarPlan
from cvar_b200.inputs import HotPathInputs

_FAMILY_BY_OWNER = {"GaussianCopulaVaR": "gaussian", "StudentCopulaVaR": "student", "PlackettCopulaVaR": "plackett"}
_PLANS = OrderedDict()      # run constants -> VarPlan (a few live plans; least recently used is closed)
_MAX_PLANS = 4


def copula_family_of(copula_density, unpack_copula_params, copula_params) -> str:
    """'gaussian' | 'student' | 'plackett' from the hooks the caller passes (see module docstring)."""
    for fn in (copula_density, unpack_copula_params):
        fam = getattr(fn, "copula_family", None) or getattr(getattr(fn, "__self__", None), "copula_family", None)
        if fam:
            return fam
        owner = getattr(fn, "__qualname__", "").split(".")[0]
        if owner in _FAMILY_BY_OWNER:
            return _FAMILY_BY_OWNER[owner]
    first, corr = unpack_copula_params(copula_params)   # last resort: the shape of what the hook unpacks
    if corr is None:
        return "plackett"                                # (theta, None)   plackett_estimation.py:12-17
    return "gaussian" if first is None else "student"    # (None, R) gaussian_estimation.py:13-24 / (nu, R) student_estimation.py:40-56


def _plan(inp: HotPathInputs, clip_lo: float) -> VarPlan:
    states = b"" if inp.sigma_states is None else inp.sigma_states.tobytes()
    key = (inp.copula, inp.marginal, int(inp.n), *(None if np.isnan(p) else float(p) for p in (inp.rho, inp.nu, inp.theta)),
           tuple(float(w) for w in inp.weights), inp.x.tobytes(), inp.dx.tobytes(), states, float(clip_lo))
    plan = _PLANS.pop(key, None)
    if plan is None:
        plan = VarPlan(inp, clip_lo=clip_lo)
        while len(_PLANS) >= _MAX_PLANS:
            _PLANS.popitem(last=False)[1].close()
    _PLANS[key] = plan
    return plan


def calc_grids_and_integrals_results(
        T,
        unique_var_values,
        unique_indices,
        num_points,
        dim,
        var_function,
        lower_bound,
        upper_bound,
        grids_generations_params,
        integrations_params_t,
        integrations_params_static,
        copula_params,
        integrated_function,
        copula_density,
        unpack_copula_params,
        weights
                                     ):
    """Strip mass of every day t in range(T) for its bounds `unique_var_values[unique_indices[t]]` = (lower, upper):
    np.ndarray[T], what the reference's function of the same name returns (calc_integral.py:8-119)."""
    _lib.check_dim(int(dim))
    _lib.check_dim(len(weights))
    family = copula_family_of(copula_density, unpack_copula_params, copula_params)
    bounds = np.asarray(unique_var_values, dtype=float).reshape(-1, 2)[np.asarray(unique_indices).reshape(-1)]
    if bounds.shape[0] != int(T):
        raise ValueError(f"unique_indices maps {bounds.shape[0]} days, T = {T}")
    _, x, dx, _ = grids_generations_params
    kw = dict(copula=family, marginal="single" if integrations_params_static is None else "mixture", n=int(num_points),
              x=x, dx=dx, weights=np.asarray(weights, float))
    first, corr = unpack_copula_params(copula_params)   # (None | nu | theta, correlation matrix | None)
    if family == "student":
        kw["nu"], kw["rho"] = float(first), float(np.asarray(corr)[0, 1])
    elif family == "gaussian":
        kw["rho"] = float(np.asarray(corr)[0, 1])
    else:
        kw["theta"] = float(first)
    if integrations_params_static is None:
        kw["sigma"] = np.asarray(integrations_params_t[0], float)[: int(T)]
    else:
        kw["probs"] = np.asarray(integrations_params_t[0], float)[: int(T)]
        kw["sigma_states"] = np.asarray(integrations_params_static, float)
    inp = HotPathInputs(**kw)
    return _plan(inp, float(lower_bound)).strip_mass(inp.day_params(), bounds)
