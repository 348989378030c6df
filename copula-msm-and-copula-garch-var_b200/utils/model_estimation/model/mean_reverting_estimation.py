"""Kalman mean-reverting log-vol marginal adapter
(mirror of utils/model_estimation/model/mean_reverting_estimation.py:11-252)."""
from utils.model_estimation.model._single_normal import SingleNormalEstimation


class MeanRevertingEstimation(SingleNormalEstimation):
    model_name = "Kalman mean-reverting"
