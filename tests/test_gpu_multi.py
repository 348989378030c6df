"""Two-GPU NCCL run of the day-sharded solve (skipped unless two CUDA devices are visible).

Each rank solves its block of days, the decision words are all-gathered, every rank finalises the whole batch;
the result must be bit-identical to the single-GPU solve of the full batch (including the batch-wide iteration
count, quirk Q7: here one shard alone would stop after 21 iterations while the full batch needs 22)."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from cvar_b200.backend import VarPlan
        from cvar_b200.distributed import shard_bounds, solve_sharded
        inp = _inputs()
        lo, hi = shard_bounds(inp.T, world, rank)
        mine = inp.take_days(slice(lo, hi))
        with VarPlan(mine, device=rank) as plan:
            d = torch.from_numpy(mine.day_params()).cuda(rank)
            var, case, iters = solve_sharded(plan, d, inp.T, [0.01, 0.05], ptf_mean=0.125)
            torch.cuda.synchronize()
            np.savez(os.path.join(out_dir, f"r{rank}.npz"), var=var.cpu().numpy(), case=case.cpu().numpy(),
                     iters=iters.cpu().numpy())
    finally:
        dist.destroy_process_group()


def _inputs():
    from cvar_b200.inputs import make_inputs
    # calm days first (bracket C / D: 21 iterations), turbulent days last (bracket A: 22 iterations)
    sigma = np.vstack([np.full((5, 2), 0.9), np.full((4, 2), 2.8)]) + np.linspace(0, 0.05, 9)[:, None]
    return make_inputs("student", "single", 128, rho=0.6, nu=5.3, sigma=sigma, ptf_mean=0.125)


def test_two_gpu_sharded_solve_equals_single_gpu(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    import torch.multiprocessing as mp
    from conftest import PKG_ROOT
    from cvar_b200.backend import VarPlan
    os.environ["PYTHONPATH"] = f"{PKG_ROOT}:{os.path.dirname(__file__)}:{os.environ.get('PYTHONPATH', '')}"
    inp = _inputs()
    with VarPlan(inp, device=0) as plan:
        full = plan.solve(inp.day_params(), [0.01, 0.05], ptf_mean=0.125)
        first_shard = plan.solve(inp.day_params()[:5], [0.01, 0.05], ptf_mean=0.125)
    assert full.iterations[0] == 22 and first_shard.iterations[0] == 21          # the coupling is really exercised
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        got = np.load(os.path.join(tmp_path, f"r{r}.npz"))
        assert got["var"].tobytes() == full.var.tobytes()
        assert np.array_equal(got["case"], full.case) and np.array_equal(got["iters"], full.iterations)


def _pipeline_case():
    from cvar_b200 import synthetic as syn
    from cvar_b200.forecast import MsmParams
    k, N, T = 5, 70, 9
    prm = [MsmParams(0.4, 1.1, 3.0, 0.3), MsmParams(0.55, 1.4, 5.0, 0.2)]
    series = np.array([syn.msm_simulate_returns(T + N - 1, k, p.m0, p.sigma_bar, p.b, p.gamma, 90 + i) for i, p in enumerate(prm)])
    return k, N, T, prm, series


def _pipeline_worker(rank, world, port, out_dir):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from cvar_b200 import forecast as fc
        from cvar_b200.backend import VarPlan
        from cvar_b200.distributed import var_from_returns_sharded
        from cvar_b200.inputs import make_inputs
        k, N, T, prm, series = _pipeline_case()
        _, sig, _ = fc.msm_forecast(series[:, :N], prm, k, N, device=rank)          # vol levels only (run constants)
        inp = make_inputs("student", "mixture", 96, rho=0.6, nu=5.3, probs=np.full((1, 2, sig.shape[1]), 1.0 / sig.shape[1]),
                          sigma_states=sig)
        with VarPlan(inp, device=rank) as plan:
            var, case, iters = var_from_returns_sharded(plan, lambda r: fc.msm_forecast_device(r, prm, k, N)[0], series, N,
                                                        [0.01, 0.05])
            torch.cuda.synchronize()
            np.savez(os.path.join(out_dir, f"p{rank}.npz"), var=var.cpu().numpy(), case=case.cpu().numpy())
    finally:
        dist.destroy_process_group()


def test_two_gpu_returns_to_var_equals_single_gpu(tmp_path):
    """Forecast + solve sharded by window block over two GPUs == host forecast + single-GPU solve, bit for bit."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    import torch.multiprocessing as mp
    from conftest import PKG_ROOT
    from cvar_b200 import forecast as fc
    from cvar_b200.backend import VarPlan
    from cvar_b200.inputs import make_inputs
    os.environ["PYTHONPATH"] = f"{PKG_ROOT}:{os.path.dirname(__file__)}:{os.environ.get('PYTHONPATH', '')}"
    k, N, T, prm, series = _pipeline_case()
    pbs, sig, _ = fc.msm_forecast(series, prm, k, N, device=0)
    inp = make_inputs("student", "mixture", 96, rho=0.6, nu=5.3, probs=pbs, sigma_states=sig)
    with VarPlan(inp, device=0) as plan:
        full = plan.solve(inp.day_params(), [0.01, 0.05])
    mp.spawn(_pipeline_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        got = np.load(os.path.join(tmp_path, f"p{r}.npz"))
        assert got["var"].tobytes() == full.var.tobytes() and np.array_equal(got["case"], full.case)
