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
