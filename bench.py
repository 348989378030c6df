#!/usr/bin/env python
"""Benchmark of the per-day VaR solve: `python bench.py --gpus N --steps K --warmup W [--impl reference]`.

One step = one pass of the hot path over one batch: every (day, alpha) solve of the workload.  The default
workload is BASELINE.json configs[2] -- the configuration the metric is quoted on ("solves/sec at a 2048^2
grid"): Student-t copula + MSM k=8 (q = 9 merged vol states) mixture marginals, 99 % VaR, 1000 days, n = 2048.
Other configurations: --workload c1|c2|c3|c4|c5_<copula>_<single|mixture>.

Prints ONE JSON line (rank 0).  `value` is whole-job throughput with the per-day parameters resident in HBM;
`e2e` goes through the host-buffer C ABI (pinned host memory in, VaR vector out); `roofline` is the solve
kernel against the FP64-pipe peak measured in the same run; `cpu_baseline` is the NumPy/SciPy oracle port timed
on this box's host cores.  `--impl reference` times that CPU port alone (the reference itself is pure
Python/numba and does not exist on the GPU box).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent
for _p in (str(REPO / "copula-msm-and-copula-garch-var_b200"), str(REPO)):
    if _p not in sys.path:
        sys.path.insert(0, _p)

METRIC = "VaR solves/sec (day x alpha)"
UNIT = "solves/s"
DEFAULT_WORKLOAD = "c3"
# dram__bytes_read.sum + dram__bytes_write.sum of one solve_kernel launch, from the committed ncu --set full captures
# (profiles/r1_c3_solve_kernel_v12_ncu_summary.txt, profiles/r1_c4_solve_kernel_v12_ncu_summary.txt); the kernel
# reads 144 KB (c3) / 16 KB (c4) of per-day parameters plus the plan tables (c3: 0.6 MB of mixture state tables, mostly
# L2 hits) and writes 8-16 KB of decision words that stay in L2.
NCU_DRAM_BYTES_PER_LAUNCH = {"c3": 889856, "c4": 140032}
# sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active from the same captures: the ISSUED-instruction view of
# the FP64 pipe, next to the algorithmic-flop fraction (which can exceed 1)
NCU_FP64_PIPE_PCT = {"c3": 51.9, "c4": 50.0}
WORKLOAD_DESCRIPTIONS = {
    "c1": "BASELINE configs[0]: Gaussian copula + GARCH(1,1) sigma path, n=100, 99% VaR",
    "c2": "BASELINE configs[1]: Student-t copula + GARCH(1,1), n=1024, 95%/99% VaR",
    "c3": "BASELINE configs[2]: Student-t copula + MSM k=8 (q=9 mixture marginals), n=2048, 99% VaR",
    "c4": "BASELINE configs[3]: Plackett copula + Kalman mean-reverting, n=2048, 95%/99% VaR",
}


# ------------------------------------------------------------------------------------------------
# CPU port (the oracle) timed on host cores -- only used by the cpu_baseline leg and --impl reference
# ------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    os.environ["OMP_NUM_THREADS"] = os.environ["MKL_NUM_THREADS"] = os.environ["OPENBLAS_NUM_THREADS"] = "1"
    inp, alpha, days, forced = args
    from oracle import var_oracle as vo
    tr = vo.calc_var(inp, alpha, days=list(days), forced_iterations=forced)
    return list(days), tr.var


def cpu_port_solve(inp, alphas, days, forced_by_alpha, pool, workers):
    """Oracle VaR for `days` x `alphas` on `workers` processes. Returns (seconds, {alpha: var[len(days)]})."""
    chunks = [c for c in np.array_split(np.asarray(days), workers) if len(c)]
    jobs = [(inp, a, c.tolist(), None if forced_by_alpha is None else int(forced_by_alpha[k]))
            for k, a in enumerate(alphas) for c in chunks]
    t0 = time.perf_counter()
    results = list(pool.map(_cpu_worker, jobs))
    dt = time.perf_counter() - t0
    out = {}
    pos = {d: i for i, d in enumerate(days)}
    it = iter(results)
    for a in alphas:
        v = np.empty(len(days))
        for _ in chunks:
            ds, vals = next(it)
            for d, x in zip(ds, vals):
                v[pos[d]] = x
        out[a] = v
    return dt, out


def make_pool(workers):
    import multiprocessing as mp
    from concurrent.futures import ProcessPoolExecutor
    return ProcessPoolExecutor(max_workers=workers, mp_context=mp.get_context("spawn"))


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled in the background (B200_PROFILING.md clocks line)."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int, period_ms: int = 20):
        self.samples, self.proc, self.thread = [], None, None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", str(period_ms),
                 "-i", str(index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.perf_counter(), line.strip()))

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self, t0, t1):
        rows = [s for t, s in self.samples if t0 <= t <= t1]
        window = "timed_region"
        if not rows:
            rows, window = [s for _, s in self.samples], "whole_run"
        sm, mx, pw, reasons = [], [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "window": window}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "reasons": sorted(reasons), "samples": len(sm), "window": window}


# ------------------------------------------------------------------------------------------------
def build_workload(name, days_per_gpu, n, world, rank):
    from cvar_b200 import synthetic as syn
    from cvar_b200.distributed import shard_bounds
    defaults_T = {"c1": 250, "c2": 1000, "c3": 1000, "c4": 1000}
    per = days_per_gpu or defaults_T.get(name, 1000)
    inp_all, alphas = syn.baseline_config(name, T=per * world, n=n)
    lo, hi = shard_bounds(per * world, world, rank)
    return inp_all, inp_all.take_days(slice(lo, hi)), tuple(alphas), per


def workload_config(name, inp, alphas, per, world):
    return {
        "workload": WORKLOAD_DESCRIPTIONS.get(name, f"BASELINE configs[4] member {name}"), "name": name,
        "copula": inp.copula, "marginal": inp.marginal, "q": inp.q, "grid": f"{inp.n}x{inp.n}",
        "days_per_gpu": per, "days_total": per * world, "alphas": list(alphas),
        "solves_per_step": per * world * len(alphas), "sharding": f"days over {world} GPU(s), one all-gather of decision words",
        "l2": "256 MiB scratch write between timed steps (inputs are < 1 MiB and compute-bound)",
    }


def run_reference(args, world, rank):
    """--impl reference: the CPU port of the path on all host cores, bounded sample of the same workload."""
    if rank != 0:
        return
    name = args.workload
    inp_all, _, alphas, per = build_workload(name, args.days, args.n, 1, 0)
    workers = len(os.sched_getaffinity(0))
    sample = list(range(0, inp_all.T, max(1, inp_all.T // args.cpu_sample_days)))[: args.cpu_sample_days]
    pool = make_pool(workers)
    try:
        cpu_port_solve(inp_all, alphas[:1], sample[:workers], None, pool, workers)     # spawn + import the workers (untimed)
        for _ in range(args.warmup):
            cpu_port_solve(inp_all, alphas[:1], sample[:workers], None, pool, workers)
        t = 0.0
        for _ in range(args.steps):
            dt, _ = cpu_port_solve(inp_all, alphas, sample, None, pool, workers)
            t += dt
    finally:
        pool.shutdown()
    nsolve = len(sample) * len(alphas)
    value = nsolve * args.steps / t
    sample_txt = f"{len(sample)} of {inp_all.T} days x {len(alphas)} alpha(s) per step, n={inp_all.n}"
    cfg = workload_config(name, inp_all, alphas, per, 1)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": workers, "kind": "port", "sample": sample_txt},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "NumPy/SciPy restatement of the reference's strip scheme (oracle/var_oracle.py), one process per host "
                "core; the reference itself (numba/joblib, materialised grids) is slower and cannot hold n=2048 mixtures",
    }))


def run_b200(args, world, rank, local_rank):
    import torch
    import torch.distributed as dist
    from cvar_b200.backend import VarPlan, fp64_peak_tflops
    from cvar_b200.distributed import solve_sharded
    from cvar_b200.workmodel import algorithmic_flops

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the VaR backend has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    name = args.workload
    inp_all, inp, alphas, per = build_workload(name, args.days, args.n, world, rank)
    T_total, na = inp_all.T, len(alphas)
    plan = VarPlan(inp, device=local_rank)
    info = plan.info()

    d_day = torch.from_numpy(inp.day_params()).to(dev)
    scratch = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    sampler = ClockSampler(local_rank) if rank == 0 else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        return solve_sharded(plan, d_day, T_total, alphas, ptf_mean=inp.ptf_mean)

    for _ in range(args.warmup):
        var, case, iters = step()
    barrier()
    # ---- value: device-resident inputs, K steps, CUDA events, max over ranks -------------------------
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t_region0 = time.perf_counter()
    for a, b in ev:
        scratch.zero_()                      # L2 flush, outside the per-step timing
        a.record()
        var, case, iters = step()
        b.record()
    barrier()
    t_region1 = time.perf_counter()
    total_ms = sum(a.elapsed_time(b) for a, b in ev)
    tt = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    total_ms = float(tt.item())
    value = T_total * na * args.steps / (total_ms * 1e-3)

    # ---- roofline: the solve kernel alone, on the launching stream ------------------------------------
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    traj = torch.empty((na, inp.T, 2), dtype=torch.int32, device=dev)
    for a, b in kev:
        scratch.zero_()
        a.record()
        plan.solve_device(d_day, alphas, traj=traj)
        b.record()
    torch.cuda.synchronize()
    kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in kev]))
    cells = torch.zeros((na, inp.T), dtype=torch.int64, device=dev)
    plan.solve_device(d_day, alphas, traj=traj, cells=cells)
    torch.cuda.synchronize()
    cells_np = cells.cpu().numpy()
    flops_launch = algorithmic_flops(inp.copula, inp.marginal, inp.q, inp.n, cells_np)
    achieved_tf = flops_launch / (kernel_ms * 1e-3) / 1e12
    peak_tf, peak_ms = fp64_peak_tflops(local_rank, 100.0)

    # ---- e2e: host buffers through the public API, H2D + kernels + D2H inside the timed region ---------
    pin_in = torch.from_numpy(np.ascontiguousarray(inp.day_params())).pin_memory()
    pin_out = torch.empty((na, T_total if world > 1 else inp.T), dtype=torch.float64).pin_memory()
    if world == 1:
        day_np, out_np = pin_in.numpy(), pin_out.numpy()

        def e2e_step():
            plan.solve(day_np, alphas, ptf_mean=inp.ptf_mean, out=out_np)
    else:
        d_in = torch.empty_like(d_day)

        def e2e_step():
            d_in.copy_(pin_in, non_blocking=True)
            v, _, _ = solve_sharded(plan, d_in, T_total, alphas, ptf_mean=inp.ptf_mean)
            pin_out.copy_(v, non_blocking=True)
            torch.cuda.synchronize()
    for _ in range(min(args.warmup, 3)):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = T_total * na * args.steps / float(te.item())
    h2d = int(pin_in.numel() * 8) * world
    d2h = int(pin_out.numel() * 8) * (world if world > 1 else 1)

    if sampler:
        sampler.stop()
    if rank != 0:
        return
    clocks = sampler.summary(t_region0, t_region1)

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": workload_config(name, inp_all, alphas, per, world),
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "path": "cvar_solve_host (pinned host buffers)" if world == 1 else
                        "pinned H2D + cvar_solve_device + NCCL all-gather + cvar_finalize_device + D2H"},
        # kernels of this repo per timed step: [order_key_kernel when the batch exceeds two waves of CTAs,] solve_kernel,
        # finalize_reduce_kernel, finalize_apply_kernel (the radix sort of the launch order is cub's, not counted)
        "gpu_launches": (4 if inp.T > 2 * info.sm_count * max(info.ctas_per_sm, 1) else 3) * args.steps,
        "roofline": {
            "bound": "fp64", "kernel": f"solve_kernel<{info.kernel_variant}> ({inp.copula})", "achieved": achieved_tf, "peak": peak_tf,
            "unit": "TFLOP/s", "frac": achieved_tf / peak_tf, "traffic": NCU_DRAM_BYTES_PER_LAUNCH.get(name),
            "kernel_ms": kernel_ms, "algorithmic_flops_per_launch": flops_launch,
            "ncu_fp64_pipe_pct_of_active_cycles": NCU_FP64_PIPE_PCT.get(name),
            "cells_per_solve_mean": float(cells_np.mean()),
            "peak_source": f"measured in this run: dependency-free DFMA micro-benchmark, {peak_ms:.0f} ms "
                           "(MEASURED_PEAKS.json has no FP64 entry; nominal 148 SM x 64 lanes x 2 x 1.965 GHz = 37.2)",
        },
        "iterations": [int(k) for k in iters.cpu().numpy()],
        "plan": {"ctas_per_sm": info.ctas_per_sm, "threads_per_cta": info.threads_per_cta,
                 "smem_bytes_per_cta": info.smem_bytes_per_cta, "sm_count": info.sm_count,
                 "kernel_variant": info.kernel_variant},
    }

    # ---- cpu_baseline + parity on a bounded sample (rank 0, N = 1 only) ---------------------------------
    if world == 1 and args.cpu_sample_days > 0:
        from cvar_b200.backtest import exceedances      # the oracle itself is only executed inside the worker processes
        workers = len(os.sched_getaffinity(0))
        sample = list(range(0, inp.T, max(1, inp.T // args.cpu_sample_days)))[: args.cpu_sample_days]
        forced = [int(k) for k in iters.cpu().numpy()]
        pool = make_pool(workers)
        try:
            cpu_port_solve(inp, alphas[:1], sample[:workers], forced[:1], pool, workers)    # spin the workers up
            dt, cpu_var = cpu_port_solve(inp, alphas, sample, forced, pool, workers)
        finally:
            pool.shutdown()
        gpu_var = var.cpu().numpy()
        max_dvar = max(float(np.max(np.abs(gpu_var[k][sample] - cpu_var[a]))) for k, a in enumerate(alphas))
        rng = np.random.default_rng(11)
        r_ptf = rng.standard_normal(len(sample)) * 1.2
        exc_equal = all(exceedances(gpu_var[k][sample], r_ptf) == exceedances(cpu_var[a], r_ptf)
                        for k, a in enumerate(alphas))
        out["cpu_baseline"] = {"value": len(sample) * na / dt, "unit": UNIT, "cores": workers, "kind": "port",
                               "sample": f"{len(sample)} of {inp.T} days x {na} alpha(s), n={inp.n}, {dt:.1f} s wall"}
        out["parity"] = {"max_abs_dvar_vs_oracle": max_dvar, "exceedance_counts_equal": bool(exc_equal),
                         "days_checked": len(sample)}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD)
    ap.add_argument("--days", type=int, default=None, help="days per GPU (default: the configuration's own T)")
    ap.add_argument("--n", type=int, default=None, help="grid points per axis (default: the configuration's own n)")
    ap.add_argument("--cpu-sample-days", type=int, default=128,
                    help="days of the workload the CPU port solves (cpu_baseline leg / one --impl reference step)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, world, rank)
        return
    if world > 1:
        import torch
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_b200(args, world, rank, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
