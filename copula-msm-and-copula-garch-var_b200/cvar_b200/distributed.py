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


def gather_blocks(traj_local: torch.Tensor, T_total: int, group=None, out: torch.Tensor | None = None) -> torch.Tensor:
    """All-gather the (n_alpha, T_local, 2) blocks into (world, n_alpha, per, 2), per = ceil(T_total / world): the
    collective's own output layout, which `cvar_finalize_blocked_device` reads as it is -- no concatenate / permute /
    copy kernel on either side of the collective.  Only a ragged last block is padded (one small copy on that rank)."""
    world = dist.get_world_size(group)
    na, t_local, two = traj_local.shape
    per = -(-T_total // world)
    if t_local > per:
        raise ValueError(f"local block of {t_local} days exceeds ceil(T/world) = {per}")
    if t_local < per:
        padded = torch.zeros((na, per, two), dtype=traj_local.dtype, device=traj_local.device)
        padded[:, :t_local] = traj_local
        traj_local = padded
    if out is None:
        out = torch.empty((world, na, per, two), dtype=traj_local.dtype, device=traj_local.device)
    dist.all_gather_into_tensor(out.view(-1), traj_local.contiguous().view(-1), group=group)
    return out


def solve_sharded(plan, day_params_local: torch.Tensor, T_total: int, alphas, ptf_mean: float = 0.0, group=None):
    """Solve this rank's days, gather the decision words, finalize the whole batch on every rank.

    Returns (var[n_alpha, T_total], case[n_alpha, T_total], iterations[n_alpha]) as CUDA tensors.
    """
    traj_local = plan.solve_device(day_params_local, alphas)
    if not (dist.is_initialized() and dist.get_world_size(group) > 1):
        return plan.finalize_device(traj_local, ptf_mean=ptf_mean)
    return plan.finalize_device(gather_blocks(traj_local, T_total, group), ptf_mean=ptf_mean, T=T_total)


class ShardedSolver:
    """Repeated sharded solves of batches of one shape, two batches in flight.

    Consecutive steps alternate between TWO streams, each with its own plan (a plan owns launch-order scratch and
    iteration counters, so overlapping launches need one plan each) and its own buffers.  A step is solve kernel ->
    all-gather -> finalize on its stream; while it drains, the next step's solve kernel is already running on the other
    stream.  That hides the two things a lone batch pays for: the collective's latency (~20-30 us on 8 GPUs) and the
    partly idle GPU at the end of every solve launch (a 1000-day launch is 3.4 waves of one-CTA days: SMs 93 % active).
    Results of `step` are valid on the step's stream: call `synchronize()` (the caller's stream waits for both) before
    reading them.  With a single plan the steps still alternate buffers but share one stream for the solves (only the
    collective + finalize overlap).  `phase_us()` times the three phases of one step separately for the benchmark record.
    """

    def __init__(self, plan, T_total: int, n_alpha: int, t_local: int, group=None, second_plan=None):
        self.plans = [plan, second_plan if second_plan is not None else plan]
        self.plan, self.T, self.na, self.group = plan, int(T_total), int(n_alpha), group
        self.sharded = dist.is_initialized() and dist.get_world_size(group) > 1
        self.world = dist.get_world_size(group) if self.sharded else 1
        self.per = -(-self.T // self.world)
        dev = torch.device("cuda", plan.device)
        self.two_streams = second_plan is not None
        self.side = torch.cuda.Stream(device=dev)
        self.streams = [torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)] if self.two_streams else None
        mk = lambda *shape, dtype: torch.empty(shape, dtype=dtype, device=dev)     # noqa: E731
        self.buffers = []
        for _ in range(2):
            local = torch.zeros((self.na, self.per if self.sharded else t_local, 2), dtype=torch.int32, device=dev)
            self.buffers.append(dict(
                local=local, view=local[:, :t_local] if t_local < local.shape[1] else local,
                blocks=mk(self.world, self.na, self.per, 2, dtype=torch.int32) if self.sharded else None,
                var=mk(self.na, self.T, dtype=torch.float64), case=mk(self.na, self.T, dtype=torch.int32),
                iters=mk(self.na, dtype=torch.int32), free=None))
        self.t_local, self.flip = int(t_local), 0
        for pl in set(self.plans):
            pl.reserve(max(self.t_local, 1))

    def _solve(self, plan, buf, day_local, alphas):
        # a ragged last block solves into a (n_alpha, t_local, 2) scratch and is copied into the padded block
        if buf["view"] is buf["local"]:
            plan.solve_device(day_local, alphas, traj=buf["local"])
        else:
            buf["local"][:, : self.t_local] = plan.solve_device(day_local, alphas)

    def _finish(self, plan, buf, ptf_mean):
        if self.sharded:
            dist.all_gather_into_tensor(buf["blocks"].view(-1), buf["local"].view(-1), group=self.group)
            plan.finalize_device(buf["blocks"], ptf_mean=ptf_mean, var=buf["var"], case=buf["case"],
                                 iterations=buf["iters"], T=self.T)
        else:
            plan.finalize_device(buf["local"], ptf_mean=ptf_mean, var=buf["var"], case=buf["case"], iterations=buf["iters"])

    def step(self, day_local: torch.Tensor, alphas, ptf_mean: float = 0.0):
        k = self.flip
        buf, plan = self.buffers[k], self.plans[k]
        self.flip ^= 1
        main = torch.cuda.current_stream(self.side.device)
        if self.two_streams:
            ready = torch.cuda.Event()
            ready.record(main)                    # the step's inputs are ready on the caller's stream
            with torch.cuda.stream(self.streams[k]):
                self.streams[k].wait_event(ready)
                self._solve(plan, buf, day_local, alphas)      # stream order covers the reuse of this buffer set
                self._finish(plan, buf, ptf_mean)
            return buf["var"], buf["case"], buf["iters"]
        if buf["free"] is not None:
            main.wait_event(buf["free"])          # the side stream is done with this buffer set (two steps ago)
        self._solve(plan, buf, day_local, alphas)
        solved = torch.cuda.Event()
        solved.record(main)
        with torch.cuda.stream(self.side):
            self.side.wait_event(solved)
            self._finish(plan, buf, ptf_mean)
            buf["free"] = torch.cuda.Event()
            buf["free"].record(self.side)
        return buf["var"], buf["case"], buf["iters"]

    def synchronize(self):
        main = torch.cuda.current_stream(self.side.device)
        main.wait_stream(self.side)
        if self.two_streams:
            for st in self.streams:
                main.wait_stream(st)

    def phase_us(self, day_local: torch.Tensor, alphas, ptf_mean: float = 0.0, repeats: int = 5) -> dict:
        """{'solve': us, 'gather': us, 'finalize': us}: one step's phases one after the other on the current stream."""
        buf = self.buffers[0]
        self.synchronize()
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        acc = [0.0, 0.0, 0.0]
        for _ in range(repeats):
            if self.sharded:
                dist.barrier(group=self.group)
            ev[0].record()
            self._solve(self.plan, buf, day_local, alphas)
            ev[1].record()
            if self.sharded:
                dist.all_gather_into_tensor(buf["blocks"].view(-1), buf["local"].view(-1), group=self.group)
            ev[2].record()
            src = buf["blocks"] if self.sharded else buf["local"]
            self.plan.finalize_device(src, ptf_mean=ptf_mean, var=buf["var"], case=buf["case"], iterations=buf["iters"],
                                      T=self.T if self.sharded else None)
            ev[3].record()
            torch.cuda.synchronize()
            for k in range(3):
                acc[k] += ev[k].elapsed_time(ev[k + 1]) * 1e3 / repeats
        return {"solve": acc[0], "gather": acc[1], "finalize": acc[2]}


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
