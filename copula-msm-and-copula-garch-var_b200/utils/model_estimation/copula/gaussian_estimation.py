"""Gaussian-copula calculator (mirror of utils/model_estimation/copula/gaussian_estimation.py:7-79)."""
import numpy as np

from utils.model_estimation.copula._base import CopulaVaRBase, corr_from_rho


class GaussianCopulaVaR(CopulaVaRBase):
    copula_family = "gaussian"

    @staticmethod
    def unpack_copula_params(copula_params):
        """[rho...] -> (None, correlation matrix)."""
        return None, corr_from_rho(copula_params)

    @staticmethod
    def copula_integrations_params(best_g_params):
        corr = best_g_params["corr_matrix"]
        return corr[np.triu_indices_from(corr, k=1)]

    @staticmethod
    def copula_density(cdf, corr_matrix, **kwargs):
        from cvar_b200.density import copula_density_gpu
        return copula_density_gpu("gaussian", cdf, rho=float(np.asarray(corr_matrix)[0, 1]))
