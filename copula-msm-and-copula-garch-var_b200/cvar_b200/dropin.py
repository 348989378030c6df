"""Patch the REFERENCE's own `ValueAtRiskCalcualtion` so that its hot path runs on the B200 backend.

Use when the reference checkout is on sys.path (its `utils` package, its fitters, its data loader) and only
the per-day solve should move to the GPU:

    import cvar_b200.dropin as dropin
    from utils.calc_var_class import ValueAtRiskCalcualtion      # the reference's class
    dropin.install(ValueAtRiskCalcualtion)
    ...
    var = calculator.calc_var(obj_var=0.01)                      # now one CUDA launch

`install` replaces `calc_var` and `compute_integral` (reference: utils/calc_var_class.py:95-212); nothing else
of the reference is touched.  The GPU inputs are read from the same attributes the reference's methods read.
`install_factory` adds the `backend=` keyword to the reference's `create_var_calculator` (utils/factory.py:9-31):
"b200" or "reference" per calculator, CVAR_BACKEND as the process-wide default.
"""
from __future__ import annotations

import numpy as np

from . import _lib
from .backend import VarPlan
from .inputs import HotPathInputs

_FAMILY_BY_CLASS = {"GaussianCopulaVaR": "gaussian", "StudentCopulaVaR": "student", "PlackettCopulaVaR": "plackett"}


def inputs_from_reference_object(v) -> HotPathInputs:
    """HotPathInputs from a (reference) driver object's attributes."""
    copula = getattr(v.VaRCalculationMethod, "copula_family", None) or _FAMILY_BY_CLASS[type(v.VaRCalculationMethod).__name__]
    marginal = "single" if v.integrations_params_static is None else "mixture"
    _, x, dx, _ = v.grids_generations_params
    kw = dict(copula=copula, marginal=marginal, n=int(v.num_points), x=x, dx=dx, weights=np.asarray(v.weights, float),
              ptf_mean=float(v.ptf_mean))
    cp = v.copula_params
    if copula == "gaussian":
        kw["rho"] = float(np.atleast_1d(cp)[0])
    elif copula == "student":
        kw["nu"], kw["rho"] = float(cp[0]), float(cp[1])
    else:
        kw["theta"] = float(cp)
    if marginal == "single":
        kw["sigma"] = np.asarray(v.integrations_params_t[0], float)
    else:
        kw["probs"] = np.asarray(v.integrations_params_t[0], float)
        kw["sigma_states"] = np.asarray(v.integrations_params_static, float)
    _lib.check_dim(int(getattr(v, "dim", 2)))
    return HotPathInputs(**kw)


def _plan_for(v, inp: HotPathInputs, first_guess, second_guess) -> VarPlan:
    """One plan per set of run constants: a refit that changes the copula parameters, the grid, the weights or the
    vol levels of the object gets a new plan instead of a stale cached one."""
    cache = v.__dict__.setdefault("_cvar_b200_plans", {})
    states = b"" if inp.sigma_states is None else np.ascontiguousarray(inp.sigma_states, dtype=float).tobytes()
    key = (float(first_guess), float(second_guess[0]), float(second_guess[1]), inp.copula, inp.marginal, int(inp.n),
           *(None if np.isnan(p) else float(p) for p in (inp.rho, inp.nu, inp.theta)),   # unused parameters are NaN
           tuple(float(w) for w in inp.weights),
           np.ascontiguousarray(inp.x, dtype=float).tobytes(), np.ascontiguousarray(inp.dx, dtype=float).tobytes(), states)
    if key not in cache:
        for stale in cache.values():
            stale.close()
        cache.clear()
        cache[key] = VarPlan(inp, first_guess=key[0], second_guess=key[1:3])
    return cache[key]


def backend_of(v) -> str:
    """Backend of a driver object: the `.backend` its calculator got from the factory (install_factory), else the
    environment variable CVAR_BACKEND, else "b200" (an installed drop-in defaults to the GPU)."""
    import os

    name = getattr(v.VaRCalculationMethod, "backend", None) or os.environ.get("CVAR_BACKEND") or "b200"
    name = str(name).lower()
    if name not in ("b200", "reference", "cpu"):
        raise ValueError(f"Unsupported backend {name!r} (expected 'b200' or 'reference')")
    return "reference" if name == "cpu" else name


def calc_var(self, obj_var=0.05, first_guess=-3, second_guess=(-3.5, -2)):
    if backend_of(self) == "reference":
        return type(self)._cvar_b200_originals["calc_var"](self, obj_var, first_guess, second_guess)
    inp = inputs_from_reference_object(self)
    res = _plan_for(self, inp, first_guess, second_guess).solve(inp.day_params(), [obj_var], ptf_mean=self.ptf_mean)
    return res.var[0]


def compute_integral(self, bounds):
    if backend_of(self) == "reference":
        return type(self)._cvar_b200_originals["compute_integral"](self, bounds)
    inp = inputs_from_reference_object(self)
    return _plan_for(self, inp, -3, (-3.5, -2)).strip_mass(inp.day_params(), np.asarray(bounds, float))


def install(cls):
    """Replace the hot-path methods of the reference's class; returns the originals for `uninstall`.

    Which path an object then takes is decided per object by `backend_of`: the factory keyword / CVAR_BACKEND."""
    if "_cvar_b200_originals" in cls.__dict__:
        return cls._cvar_b200_originals
    originals = {"calc_var": cls.calc_var, "compute_integral": cls.compute_integral}
    cls.calc_var = calc_var
    cls.compute_integral = compute_integral
    cls._cvar_b200_originals = originals
    return originals


def uninstall(cls):
    for name, fn in cls.__dict__.get("_cvar_b200_originals", {}).items():
        setattr(cls, name, fn)
    if "_cvar_b200_originals" in cls.__dict__:
        del cls._cvar_b200_originals


def install_factory(factory_cls):
    """Give the REFERENCE's factory the `backend=` keyword (utils/factory.py:9-31 takes two arguments):

        ValueAtRiskCalculationFactory.create_var_calculator("student", "msm", backend="b200")

    The calculator is the reference's own object; `.backend` ("b200" or "reference"; None = CVAR_BACKEND, default
    "b200") tells the patched driver methods (install) which path to take."""
    if "_cvar_b200_create" in factory_cls.__dict__:
        return
    original = factory_cls.create_var_calculator

    def create_var_calculator(copula_type, estimation_type, backend=None):
        calculator = original(copula_type, estimation_type)
        if backend is not None:
            if str(backend).lower() not in ("b200", "reference", "cpu"):
                raise ValueError(f"Unsupported backend {backend!r} (expected 'b200' or 'reference')")
            calculator.backend = str(backend).lower()
        return calculator

    factory_cls._cvar_b200_create = original
    factory_cls.create_var_calculator = staticmethod(create_var_calculator)


def uninstall_factory(factory_cls):
    if "_cvar_b200_create" in factory_cls.__dict__:
        factory_cls.create_var_calculator = staticmethod(factory_cls._cvar_b200_create)
        del factory_cls._cvar_b200_create
