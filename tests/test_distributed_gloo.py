"""world_size-2 gloo test of the multi-GPU plumbing (runs on CPU): trajectory all-gather with ragged shards."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, T, na, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from cvar_b200.distributed import gather_blocks, gather_trajectories, shard_bounds
        full = torch.arange(na * T * 2, dtype=torch.int32).reshape(na, T, 2)
        lo, hi = shard_bounds(T, world, rank)
        got = gather_trajectories(full[:, lo:hi, :].contiguous(), T)
        torch.save(got, os.path.join(out_dir, f"r{rank}.pt"))
        # the blocked layout the finalize kernels read without reshuffling: day d = blocks[d // per, :, d % per]
        blocks = gather_blocks(full[:, lo:hi, :].contiguous(), T)
        per = -(-T // world)
        assert blocks.shape == (world, na, per, 2)
        days = torch.arange(T)
        assert torch.equal(blocks[days // per, :, days % per].permute(1, 0, 2), full)
    finally:
        dist.destroy_process_group()


def _run(T, na, tmp_path, world=2):
    import sys
    from conftest import PKG_ROOT
    os.environ["PYTHONPATH"] = f"{PKG_ROOT}:{os.environ.get('PYTHONPATH', '')}"
    port = _free_port()
    mp.spawn(_worker, args=(world, port, T, na, str(tmp_path)), nprocs=world, join=True)
    want = torch.arange(na * T * 2, dtype=torch.int32).reshape(na, T, 2)
    for r in range(world):
        got = torch.load(os.path.join(tmp_path, f"r{r}.pt"))
        assert got.shape == want.shape and torch.equal(got, want)


def test_gather_even_shards(tmp_path):
    _run(T=10, na=2, tmp_path=tmp_path)


def test_gather_ragged_shards(tmp_path):
    _run(T=7, na=1, tmp_path=tmp_path)


def test_gather_three_ranks_ragged(tmp_path):
    _run(T=8, na=2, tmp_path=tmp_path, world=3)
