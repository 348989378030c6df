"""Plackett-copula calculator (mirror of utils/model_estimation/copula/plackett_estimation.py:6-71)."""
from utils.model_estimation.copula._base import CopulaVaRBase


class PlackettCopulaVaR(CopulaVaRBase):
    copula_family = "plackett"

    @staticmethod
    def unpack_copula_params(copula_params):
        """theta -> (theta, None): theta travels in the `nu` slot, as in the reference."""
        return copula_params, None

    @staticmethod
    def copula_integrations_params(best_p_params):
        return best_p_params["theta"]

    @staticmethod
    def copula_density(cdf, nu, **kwargs):
        from cvar_b200.density import copula_density_gpu
        return copula_density_gpu("plackett", cdf, theta=float(nu))
