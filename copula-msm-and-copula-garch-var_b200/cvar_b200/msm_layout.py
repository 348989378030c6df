"""Input layout of the MSM (normal-mixture) marginals on the hot path.

Pure array packing: merge the 2^k MSM states to the k+1 distinct vol levels, tabulate the state densities
on the axis, enumerate state pairs.  These are the functions of the reference's MSM adapter that define what
the solve reads (utils/model_estimation/model/msm_estimation.py:205-248, 283-330, 369-418).
"""
from __future__ import annotations

import itertools

import numpy as np

from .axis import build_axis


# ---- the binomial MSM(k) model itself (markov_switching_multifractal/calc_prob.py:72-108) -------------------------
def msm_multiplier_table(k: int, m0: float) -> np.ndarray:
    """(2^k, k) table of multiplier vectors, `itertools.product` order (component 0 varies slowest)."""
    return np.array(list(itertools.product([m0, 2.0 - m0], repeat=k)))


def msm_vol_states(k: int, m0: float, sigma_bar: float) -> np.ndarray:
    """vol_states[2^k] = sqrt(prod_i M_i) * sigma_bar."""
    return np.sqrt(np.prod(msm_multiplier_table(k, m0), axis=1)) * sigma_bar


def msm_switch_probs(k: int, b: float, gamma: float) -> np.ndarray:
    """gamma_i, i = 0..k-1: probability that component i is redrawn in one step."""
    return 1.0 - (1.0 - gamma) ** (b ** np.arange(k))


def msm_stay_probs(k: int, b: float, gamma: float) -> np.ndarray:
    """p_i = 1 - gamma_i / 2: probability that component i keeps its value over one step."""
    return 1.0 - msm_switch_probs(k, b, gamma) / 2.0


def msm_transition_matrix(k: int, m0: float, b: float, gamma: float) -> np.ndarray:
    """Dense P[i, j] = prod_c (p_c if equal else 1 - p_c); the Kronecker product of k 2x2 factors."""
    table = msm_multiplier_table(k, m0)
    p = msm_stay_probs(k, b, gamma)
    same = table[:, None, :] == table[None, :, :]
    return np.prod(np.where(same, p, 1.0 - p), axis=2)


def state_levels(vol_states, tol: float = 1e-6):
    """(level_of_state[S], sigma_levels[q]) with the reference's rounding (msm_estimation.py:228-229)."""
    rounded = np.round(np.asarray(vol_states, float) / tol) * tol
    uniq, inv = np.unique(rounded, return_inverse=True)
    return inv.astype(np.int32), uniq


def merge_states(vol_states, probs, tol: float = 1e-6):
    """Merge states of (numerically) equal volatility.

    vol_states : (dim, S);  probs : (dim, T, S)
    returns probs_by_state (T, dim, q) and sigma_states (dim, q), with the vol levels rounded to
    multiples of ``tol`` exactly as the reference does (msm_estimation.py:228-229) so that the merged
    sigma values are the same doubles.
    """
    vol_states = np.asarray(vol_states, float)
    probs = np.asarray(probs, float)
    dim = probs.shape[0]
    merged, levels = [], []
    for d in range(dim):
        rounded = np.round(vol_states[d] / tol) * tol
        uniq, inv = np.unique(rounded, return_inverse=True)
        # row-contiguous copies so every row is summed pairwise exactly like the reference's
        # per-day `forecasts_array[i, n, :][inverse_idx == idx].sum()`
        cols = [np.ascontiguousarray(probs[d][:, inv == j]).sum(axis=1) for j in range(len(uniq))]
        merged.append(np.stack(cols, axis=1))
        levels.append(uniq)
    if len({len(u) for u in levels}) != 1:
        raise ValueError("assets merge to different numbers of vol levels")
    return np.ascontiguousarray(np.stack(merged, axis=1)), np.array(levels)


def state_densities(sigma_states, num_points: int):
    """(densities[dim, q, n], x[n], dx[n]): N(x; 0, sigma_state) on the MSM axis (msm_estimation.py:283-330)."""
    sigma_states = np.asarray(sigma_states, float)
    x, dx = build_axis(num_points, "mixture")
    s = sigma_states[:, :, None]
    dens = (1 / (np.sqrt(2 * np.pi) * s)) * np.exp(-0.5 * (x[None, None, :] / s) ** 2)
    return dens, x, dx


def state_index_pairs(dim: int, q: int):
    """(q**dim, dim) table of state-index tuples, first asset slowest (msm_estimation.py:369-389)."""
    grids = np.meshgrid(*[np.arange(q)] * dim, indexing="ij")
    return np.stack(grids, axis=-1).reshape(-1, dim)


def pair_probabilities(probs_by_state):
    """(T, q**dim) joint probabilities of the state tuples in `state_index_pairs` order
    (msm_estimation.py:392-418; independent components => product of the per-asset probabilities)."""
    p = np.asarray(probs_by_state, float)
    T, dim, q = p.shape
    out = p[:, 0, :]
    for d in range(1, dim):
        out = (out[:, :, None] * p[:, d, None, :]).reshape(T, -1)
    return out
