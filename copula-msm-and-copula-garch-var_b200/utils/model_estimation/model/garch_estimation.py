"""GARCH(p,q) marginal adapter (mirror of utils/model_estimation/model/garch_estimation.py:11-251)."""
from utils.model_estimation.model._single_normal import SingleNormalEstimation


class GarchEstimation(SingleNormalEstimation):
    model_name = "GARCH"
