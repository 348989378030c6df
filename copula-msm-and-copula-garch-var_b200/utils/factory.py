"""Calculator factory (mirror of the reference's utils/factory.py:9-31)."""
from utils.model_estimation.copula.student_estimation import StudentCopulaVaR
from utils.model_estimation.copula.gaussian_estimation import GaussianCopulaVaR
from utils.model_estimation.copula.plackett_estimation import PlackettCopulaVaR
from utils.model_estimation.model.msm_estimation import MSMEstimation
from utils.model_estimation.model.garch_estimation import GarchEstimation
from utils.model_estimation.model.mean_reverting_estimation import MeanRevertingEstimation

_COPULAS = {"student": StudentCopulaVaR, "gaussian": GaussianCopulaVaR, "plackett": PlackettCopulaVaR}
_MODELS = {"msm": MSMEstimation, "garch": GarchEstimation, "mean_reverting": MeanRevertingEstimation}


class ValueAtRiskCalculationFactory:
    @staticmethod
    def create_var_calculator(copula_type, estimation_type, strict_reference_quirks=True):
        """Calculator for (copula_type, estimation_type); ValueError("Unsupported estimation type.") otherwise.

        Quirk Q11 of the reference (factory.py:22-23) is kept by default: ('gaussian', 'mean_reverting')
        yields a *Plackett* calculator.  Pass strict_reference_quirks=False to get the Gaussian one.
        """
        if copula_type not in _COPULAS or estimation_type not in _MODELS:
            raise ValueError("Unsupported estimation type.")
        copula_cls = _COPULAS[copula_type]
        if strict_reference_quirks and copula_type == "gaussian" and estimation_type == "mean_reverting":
            copula_cls = PlackettCopulaVaR
        return copula_cls(_MODELS[estimation_type]())
