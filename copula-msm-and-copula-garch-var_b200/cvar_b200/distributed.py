"""Day-sharded multi-GPU solve: one process per GPU, torch.distributed for the plumbing.

Out-of-sample days are independent, so each rank solves a contiguous block of days with no data-path
collective.  The single cross-day coupling of the reference -- every day is iterated until the slowest day
of the WHOLE batch has converged (utils/calc_var_class.py:278, quirk Q7) -- is resolved after ONE all-gather
of the per-solve decision words (8 bytes per (day, alpha), the same size as the VaR vector itself): every
rank then applies the batch-wide iteration count locally (`cvar_finalize_device`), so results do not depend
on the number of GPUs.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(T: int, world_size: int, rank: int) -> tuple[int, int]:
    """[start, stop) of the contiguous day block of `rank`: ceil(T / world) days per rank, last ranks ragged."""
    per = -(-T // world_size)
    start = min(rank * per, T)
    return start, min(start + per, T)


def gather_trajectories(traj_local: torch.Tensor, T_total: int, group=None) -> torch.Tensor:
    """All-gather the (n_alpha, T_local, 2) int32 trajectory blocks into the global (n_alpha, T_total, 2) tensor.

    Works on any backend (NCCL for CUDA tensors, gloo for CPU tensors in tests).  Ragged last blocks are padded
    to the common block length for the collective and the padding is dropped afterwards.
    """
    world = dist.get_world_size(group)
    na, t_local, two = traj_local.shape
    per = -(-T_total // world)
    if t_local > per:
        raise ValueError(f"local block of {t_local} days exceeds ceil(T/world) = {per}")
    if t_local < per:
        pad = torch.zeros((na, per - t_local, two), dtype=traj_local.dtype, device=traj_local.device)
        traj_local = torch.cat([traj_local, pad], dim=1)
    traj_local = traj_local.contiguous()
    # flat buffers: the one layout both NCCL and gloo accept for all_gather_into_tensor
    flat = torch.empty(world * na * per * two, dtype=traj_local.dtype, device=traj_local.device)
    dist.all_gather_into_tensor(flat, traj_local.view(-1), group=group)
    gathered = flat.view(world, na, per, two)
    # (world, na, per, 2) -> (na, world * per, 2), then drop the padding of the ragged tail
    out = gathered.permute(1, 0, 2, 3).reshape(na, world * per, two)
    return out[:, :T_total, :].contiguous()


def solve_sharded(plan, day_params_local: torch.Tensor, T_total: int, alphas, ptf_mean: float = 0.0, group=None):
    """Solve this rank's days, gather the decision words, finalize the whole batch on every rank.

    Returns (var[n_alpha, T_total], case[n_alpha, T_total], iterations[n_alpha]) as CUDA tensors.
    """
    traj_local = plan.solve_device(day_params_local, alphas)
    traj = gather_trajectories(traj_local, T_total, group) if dist.is_initialized() and dist.get_world_size(group) > 1 \
        else traj_local
    return plan.finalize_device(traj, ptf_mean=ptf_mean)


def window_slice(T: int, N: int, world_size: int, rank: int, window_stride: int = 1) -> tuple[int, int, int, int]:
    """(day_start, day_stop, ret_start, ret_stop): the rolling windows [day_start, day_stop) of `rank` read the
    returns [ret_start, ret_stop) of the (T-1)*window_stride + N long series (ret_stop == ret_start for an empty
    block).  Neighbouring ranks overlap by N - window_stride returns, the price of filtering without a halo
    exchange: 8.6 KB per asset at N = 1135 against a 1.6 ms filter."""
    d0, d1 = shard_bounds(T, world_size, rank)
    if d1 <= d0:
        return d0, d1, 0, 0
    return d0, d1, d0 * window_stride, (d1 - 1) * window_stride + N


def var_from_returns_sharded(plan, producer, returns, N: int, alphas, ptf_mean: float = 0.0, window_stride: int = 1,
                             group=None):
    """Returns -> forecasts -> VaR with nothing but the return series crossing PCIe.

    returns  : (n_assets, L) centred returns on the HOST (numpy or a pinned CPU tensor), L = (T-1)*window_stride + N,
               identical on every rank; each rank uploads only the slice its windows read
    producer : callable(cuda tensor (n_assets, L_local)) -> day_params CUDA tensor (T_local, 2[, q]) -- e.g.
               ``lambda r: msm_forecast_device(r, params, k, N)[0]`` or ``lambda r: garch_forecast_device(r, ..., N)``
    -> (var[n_alpha, T], case[n_alpha, T], iterations[n_alpha]) CUDA tensors, the same on every rank and for any
       number of ranks (the forecast of a window depends on its own returns only; the solve's single cross-day
       coupling is resolved by `solve_sharded`).
    """
    import numpy as np

    r = returns if isinstance(returns, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(returns, dtype=np.float64))
    if r.dim() != 2 or r.dtype != torch.float64:
        raise ValueError("returns must be a float64 (n_assets, L) array")
    L = r.shape[1]
    if L < N or (L - N) % window_stride:
        raise ValueError("series length does not match (T-1)*window_stride + N")
    T = (L - N) // window_stride + 1
    sharded = dist.is_initialized() and dist.get_world_size(group) > 1
    world, rank = (dist.get_world_size(group), dist.get_rank(group)) if sharded else (1, 0)
    d0, d1, r0, r1 = window_slice(T, N, world, rank, window_stride)
    dev = torch.device("cuda", plan.device)
    with torch.cuda.device(dev):
        if d1 > d0:
            local = r[:, r0:r1].contiguous().to(dev, non_blocking=True)
            day = producer(local)
        else:
            shape = (0, 2) if plan.marginal == "single" else (0, 2, plan.q)
            day = torch.empty(shape, dtype=torch.float64, device=dev)
        if day.shape[0] != d1 - d0:
            raise ValueError(f"producer returned {day.shape[0]} days for a block of {d1 - d0} windows")
        return solve_sharded(plan, day.contiguous(), T, alphas, ptf_mean=ptf_mean, group=group)
