"""cvar_b200 -- B200-native backend of the per-day copula VaR solve.

Light modules (axis, inputs, synthetic, msm_layout) import without CUDA; `backend` (and everything that
solves) needs the built libcvar_b200.so and a CUDA device and fails loudly otherwise.
"""
from .inputs import HotPathInputs, make_inputs  # noqa: F401

__all__ = ["HotPathInputs", "make_inputs"]
