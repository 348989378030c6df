"""Calculator factory (mirror of the reference's utils/factory.py:9-31) with the backend selection of SURVEY section 8(b).

`create_var_calculator(copula_type, estimation_type, backend=None)`: `backend` (or the environment variable
CVAR_BACKEND when it is None; default "b200") is recorded on the calculator as `.backend`.  This package only
contains the B200 path -- there is no CPU implementation to fall back to -- so any other value raises; the same
keyword on the REFERENCE's own factory (after `cvar_b200.dropin.install_factory`) switches a reference object
between its numba/joblib path ("reference") and the CUDA library ("b200") at run time.
"""
import os

from utils.model_estimation.copula.student_estimation import StudentCopulaVaR
from utils.model_estimation.copula.gaussian_estimation import GaussianCopulaVaR
from utils.model_estimation.copula.plackett_estimation import PlackettCopulaVaR
from utils.model_estimation.model.msm_estimation import MSMEstimation
from utils.model_estimation.model.garch_estimation import GarchEstimation
from utils.model_estimation.model.mean_reverting_estimation import MeanRevertingEstimation

_COPULAS = {"student": StudentCopulaVaR, "gaussian": GaussianCopulaVaR, "plackett": PlackettCopulaVaR}
_MODELS = {"msm": MSMEstimation, "garch": GarchEstimation, "mean_reverting": MeanRevertingEstimation}
BACKENDS = ("b200",)


def resolve_backend(backend=None):
    """`backend` argument > CVAR_BACKEND > "b200"; ValueError for a backend this package does not contain."""
    name = (backend or os.environ.get("CVAR_BACKEND") or "b200").lower()
    if name not in BACKENDS:
        raise ValueError(
            f"Unsupported backend {name!r}: this package is the B200 backend and has no CPU fallback "
            "(the numba/joblib path lives in the reference checkout; see INTEGRATION.md for selecting between the two "
            "on the reference's own factory)")
    return name


class ValueAtRiskCalculationFactory:
    @staticmethod
    def create_var_calculator(copula_type, estimation_type, backend=None, strict_reference_quirks=True):
        """Calculator for (copula_type, estimation_type); ValueError("Unsupported estimation type.") otherwise.

        Quirk Q11 of the reference (factory.py:22-23) is kept by default: ('gaussian', 'mean_reverting')
        yields a *Plackett* calculator.  Pass strict_reference_quirks=False to get the Gaussian one.
        """
        if copula_type not in _COPULAS or estimation_type not in _MODELS:
            raise ValueError("Unsupported estimation type.")
        name = resolve_backend(backend)
        copula_cls = _COPULAS[copula_type]
        if strict_reference_quirks and copula_type == "gaussian" and estimation_type == "mean_reverting":
            copula_cls = PlackettCopulaVaR
        calculator = copula_cls(_MODELS[estimation_type]())
        calculator.backend = name
        return calculator
