"""Student-t-copula calculator (mirror of utils/model_estimation/copula/student_estimation.py:7-91)."""
import numpy as np

from utils.model_estimation.copula._base import CopulaVaRBase, corr_from_rho


class StudentCopulaVaR(CopulaVaRBase):
    copula_family = "student"

    @staticmethod
    def copula_integrations_params(best_t_params):
        """{'optimized_params': [nu, ...], 'corr_matrix': R} -> array([nu, rho...])."""
        corr = best_t_params["corr_matrix"]
        return np.concatenate((np.array([best_t_params["optimized_params"][0]]), corr[np.triu_indices_from(corr, k=1)]))

    @staticmethod
    def unpack_copula_params(copula_params):
        """array([nu, rho...]) -> (nu, correlation matrix)."""
        return copula_params[0], corr_from_rho(copula_params[1:])

    @staticmethod
    def copula_density(cdf, nu, corr_matrix, **kwargs):
        from cvar_b200.density import copula_density_gpu
        return copula_density_gpu("student", cdf, nu=float(nu), rho=float(np.asarray(corr_matrix)[0, 1]))
