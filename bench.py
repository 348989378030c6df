#!/usr/bin/env python
"""Benchmark of the per-day VaR solve: `python bench.py --gpus N --steps K --warmup W [--impl reference]`.

One step = one pass of the hot path over one batch: every (day, alpha) solve of the workload.  The default
workload is BASELINE.json configs[2] -- the configuration the metric is quoted on ("solves/sec at a 2048^2
grid"): Student-t copula + MSM k=8 (q = 9 merged vol states) mixture marginals, 99 % VaR, 1000 days, n = 2048.
Other configurations: --workload c1|c2|c3|c4|c5_<copula>_<single|mixture>.

Prints ONE JSON line (rank 0):
  value         whole-job throughput, per-day parameters resident in HBM (weak scaling: `--days` per GPU; `--scaling
                strong`: `--days-total` split over the GPUs); every step = solve kernel, all-gather of the decision words
                (N > 1), finalize, timed per step with CUDA events, max over ranks
  e2e           the same through the host-buffer API: pinned host parameters in, VaR vector out, copies inside the timing
  roofline      the solve kernel against the FP64 pipe: algorithmic flops (SURVEY 8(d) convention) over the DFMA peak
                measured in the same run, and `cell_pipe_frac`, the pipe share of the ISSUED FP64 instructions of the cells
  strong        (default line) a second, shorter measurement with the total number of days fixed at 8000
  phases_us     solve / all-gather / finalize of one step, timed one after the other
  pipelined     throughput with two batches in flight on two streams (ShardedSolver): the next step's solve covers a
                step's collective + finalize and the partly idle end of its solve launch
  cpu_baseline  the NumPy/SciPy oracle port on this box's host cores (N = 1), plus -- when a reference install is present
                (baseline/_ref) -- the UNMODIFIED reference's calc_var timed on BASELINE configs[0]
  parity        max |dVaR| against the oracle on a sample of days spread over all shards, exceedance counts
`--impl reference` times the CPU path alone: the unmodified reference for c1 (when installed), else the oracle port
(the reference's materialised grids cannot hold the n = 2048 mixture configurations).
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
for _p in (str(REPO / "copula-msm-and-copula-garch-var_b200"), str(REPO), str(REPO / "baseline")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

METRIC = "VaR solves/sec (day x alpha)"
UNIT = "solves/s"
DEFAULT_WORKLOAD = "c3"
STRONG_DAYS_TOTAL = 8000
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


def reference_unmodified_c1(repeats=2):
    """The UNMODIFIED reference's calc_var on BASELINE configs[0] (n=100, T=250), warm call; None without an install."""
    try:
        from reference_runner import reference_root, time_reference_calc_var
    except ImportError:
        return None, "baseline/reference_runner.py missing"
    if reference_root() is None:
        return None, "no reference install on this box (baseline/install_reference.sh was not run)"
    from cvar_b200 import synthetic as syn
    inp, alphas = syn.baseline_config("c1")
    try:
        r = time_reference_calc_var(inp, alphas[0], repeats=repeats)
    except Exception as exc:  # noqa: BLE001  (reported in the record, the GPU numbers do not depend on it)
        return None, f"reference run failed: {str(exc)[-300:]}"
    warm = min(r["seconds"][1:]) if len(r["seconds"]) > 1 else r["seconds"][0]
    return {"value": inp.T / warm, "unit": UNIT, "kind": "reference", "cores": len(os.sched_getaffinity(0)),
            "workload": "c1 (Gaussian + GARCH, n=100, 250 days, alpha=1%)", "seconds_warm_call": warm,
            "seconds_first_call": r["seconds"][0],
            "path": "utils/calc_var_class.py:95-177 via numba + joblib(n_jobs=-1), unmodified, from baseline/_ref"}, r["var"]


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
def build_workload(name, days_total, n, world, rank):
    from cvar_b200 import synthetic as syn
    from cvar_b200.distributed import shard_bounds
    inp_all, alphas = syn.baseline_config(name, T=days_total, n=n)
    lo, hi = shard_bounds(inp_all.T, world, rank)
    return inp_all, inp_all.take_days(slice(lo, hi)), tuple(alphas)


def default_days(name):
    return {"c1": 250, "c2": 1000, "c3": 1000, "c4": 1000}.get(name, 1000)


def workload_config(name, inp, alphas, world, scaling):
    per = -(-inp.T // world)
    return {
        "workload": WORKLOAD_DESCRIPTIONS.get(name, f"BASELINE configs[4] member {name}"), "name": name,
        "copula": inp.copula, "marginal": inp.marginal, "q": inp.q, "grid": f"{inp.n}x{inp.n}",
        "days_per_gpu": per, "days_total": inp.T, "alphas": list(alphas), "scaling": scaling,
        "solves_per_step": inp.T * len(alphas), "sharding": f"days over {world} GPU(s), one all-gather of decision words",
        "l2": "256 MiB scratch write between timed steps (inputs are < 1 MiB and compute-bound)",
    }


def traffic_record(name, info):
    """DRAM bytes of one solve_kernel launch from the committed ncu capture of THIS kernel build, or None.

    profiles/r2_traffic.json is written by tools/ncu_summary.py --traffic from `ncu --set full` reports and keyed by
    workload; an entry is only used when the kernel variant and CTA shape it was captured with match the running plan."""
    path = REPO / "profiles" / "r2_traffic.json"
    if not path.exists():
        return None
    try:
        entry = json.loads(path.read_text()).get(name)
    except (OSError, ValueError):
        return None
    if not entry or entry.get("kernel_variant") != info.kernel_variant or entry.get("threads_per_cta") != info.threads_per_cta:
        return None
    return entry


def run_reference(args, world, rank):
    """--impl reference: the reference's CPU implementation of the path on all host cores.

    c1 with a reference install: the UNMODIFIED reference (kind "reference"), one step = its calc_var over the 250 days.
    Otherwise the oracle port (kind "port") on a bounded sample of the same workload."""
    if rank != 0:
        return
    name = args.workload
    days_total = args.days_total or (args.days or default_days(name))
    inp_all, _, alphas = build_workload(name, days_total, args.n, 1, 0)
    workers = len(os.sched_getaffinity(0))
    cfg = workload_config(name, inp_all, alphas, 1, "weak")
    if name == "c1" and args.n is None and args.days is None and args.days_total is None:
        ref, _ = reference_unmodified_c1(repeats=max(2, args.warmup + args.steps))
        if ref is not None:
            value = ref["value"]
            print(json.dumps({
                "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": 1e3 * ref["seconds_warm_call"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
                "cpu_baseline": {"value": value, "unit": UNIT, "cores": workers, "kind": "reference",
                                 "sample": "all 250 days x 1 alpha, warm call"},
                "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0, "note": ref["path"]}))
            return
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
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": workers, "kind": "port", "sample": sample_txt},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "kind=port: NumPy/SciPy restatement of the reference's strip scheme (oracle/var_oracle.py), one process per "
                "host core -- a STAND-IN that is faster than the reference itself (numba/joblib, materialised grids), which "
                "cannot hold n=2048 mixtures; the unmodified reference is timed on c1 (`--impl reference --workload c1`, "
                "and `cpu_baseline.reference_unmodified` of the default line)",
    }))


def timed_steps(step, steps, scratch, barrier, dev, world):
    """K steps, each between a pair of CUDA events on the launching stream (L2 flushed before each, outside the pair);
    returns the sum of the per-step device times in ms, max over ranks."""
    import torch
    import torch.distributed as dist
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    out = None
    for a, b in ev:
        scratch.zero_()
        a.record()
        out = step()
        b.record()
    barrier()
    total_ms = sum(a.elapsed_time(b) for a, b in ev)
    tt = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    return float(tt.item()), out


def run_b200(args, world, rank, local_rank):
    import torch
    import torch.distributed as dist
    from cvar_b200.backend import VarPlan, fp64_peak_tflops
    from cvar_b200.distributed import ShardedSolver, solve_sharded
    from cvar_b200.workmodel import algorithmic_flops, cell_fp64_pipe_fraction, issued_fp64_per_cell

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the VaR backend has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    name = args.workload
    strong = args.scaling == "strong"
    days_total = (args.days_total or STRONG_DAYS_TOTAL) if strong else (args.days or default_days(name)) * world
    inp_all, inp, alphas = build_workload(name, days_total, args.n, world, rank)
    T_total, na = inp_all.T, len(alphas)
    plan = VarPlan(inp, device=local_rank)
    plan.reserve(max(inp.T, 1))
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
    t_region0 = time.perf_counter()
    total_ms, (var, case, iters) = timed_steps(step, args.steps, scratch, barrier, dev, world)
    t_region1 = time.perf_counter()
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
    plan.evaluated_cells(reset=True)
    plan.solve_device(d_day, alphas, traj=traj, cells=cells)
    torch.cuda.synchronize()
    cells_evaluated = plan.evaluated_cells(reset=True)     # strips shared between the alphas of a day are evaluated once
    cells_np = cells.cpu().numpy()
    flops_launch = algorithmic_flops(inp.copula, inp.marginal, inp.q, inp.n, cells_np)
    achieved_tf = flops_launch / (kernel_ms * 1e-3) / 1e12
    peak_tf, peak_ms = fp64_peak_tflops(local_rank, 100.0)
    cell_frac = cell_fp64_pipe_fraction(info.kernel_variant, info.pow_octaves, cells_evaluated, kernel_ms * 1e-3, peak_tf)

    # ---- phases and the pipelined variant ---------------------------------------------------------------
    plan2 = VarPlan(inp, device=local_rank)        # two batches in flight need a plan (scratch) each
    solver = ShardedSolver(plan, T_total, na, inp.T, second_plan=plan2)
    phases = solver.phase_us(d_day, alphas, ptf_mean=inp.ptf_mean, repeats=5)
    for _ in range(3):
        solver.step(d_day, alphas, ptf_mean=inp.ptf_mean)
    solver.synchronize()
    barrier()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    for _ in range(args.steps):
        solver.step(d_day, alphas, ptf_mean=inp.ptf_mean)
    solver.synchronize()
    p1.record()
    barrier()
    tp = torch.tensor([p0.elapsed_time(p1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tp, op=dist.ReduceOp.MAX)
    pipelined_value = T_total * na * args.steps / (float(tp.item()) * 1e-3)

    # ---- e2e: host buffers through the public API, H2D + kernels + D2H inside the timed region ---------
    pin_in = torch.from_numpy(np.ascontiguousarray(inp.day_params())).pin_memory()
    pin_out = torch.empty((na, T_total if world > 1 else inp.T), dtype=torch.float64).pin_memory()
    if world == 1:
        day_np, out_np = pin_in.numpy(), pin_out.numpy()

        def e2e_step():
            plan.solve(day_np, alphas, ptf_mean=inp.ptf_mean, out=out_np, details=False)
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

    # ---- strong scaling at a fixed total (default line only): 8000 days split over the ranks ------------
    strong_rec = None
    if not strong and args.strong_steps > 0 and name in ("c2", "c3", "c4"):
        s_all, s_inp, _ = build_workload(name, STRONG_DAYS_TOTAL, args.n, world, rank)
        s_plan = VarPlan(s_inp, device=local_rank)
        s_plan.reserve(max(s_inp.T, 1))
        s_day = torch.from_numpy(s_inp.day_params()).to(dev)

        def s_step():
            return solve_sharded(s_plan, s_day, s_all.T, alphas, ptf_mean=s_inp.ptf_mean)
        for _ in range(2):
            s_step()
        barrier()
        s_ms, _ = timed_steps(s_step, args.strong_steps, scratch, barrier, dev, world)
        strong_rec = {"days_total": s_all.T, "days_per_gpu": -(-s_all.T // world), "steps": args.strong_steps,
                      "ms_per_step": s_ms / args.strong_steps, "value": s_all.T * na * args.strong_steps / (s_ms * 1e-3),
                      "unit": UNIT, "note": "same workload with the total number of days fixed (strong scaling over --gpus)"}
        s_plan.close()

    if sampler:
        sampler.stop()
    var_np = var.cpu().numpy()
    iters_np = [int(k) for k in iters.cpu().numpy()]
    if rank != 0:
        return
    clocks = sampler.summary(t_region0, t_region1)
    traffic = traffic_record(name, info)
    ordered = inp.T > 2 * info.sm_count * max(info.ctas_per_sm, 1)
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": workload_config(name, inp_all, alphas, world, args.scaling),
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "path": "cvar_solve_host (pinned host buffers)" if world == 1 else
                        "pinned H2D + cvar_solve_device + NCCL all-gather + cvar_finalize_blocked_device + D2H"},
        # kernels of this repo per timed step: [order_key_kernel when the batch exceeds two waves of CTAs,] solve_kernel,
        # finalize_reduce_kernel, finalize_apply_kernel (the radix sort of the launch order is cub's, not counted)
        "gpu_launches": (4 if ordered else 3) * args.steps,
        "roofline": {
            "bound": "fp64", "kernel": f"solve_kernel<{info.kernel_variant}> ({inp.copula})", "achieved": achieved_tf, "peak": peak_tf,
            "unit": "TFLOP/s", "frac": achieved_tf / peak_tf,
            "traffic": None if traffic is None else traffic["dram_bytes"],
            "traffic_source": None if traffic is None else traffic.get("source"),
            "kernel_ms": kernel_ms, "algorithmic_flops_per_launch": flops_launch,
            "cell_pipe_frac": cell_frac,
            "issued_fp64_per_cell": issued_fp64_per_cell(info.kernel_variant, info.pow_octaves),
            "cells_per_solve_mean": float(cells_np.mean()), "cells_evaluated_per_launch": int(cells_evaluated),
            "frac_note": "frac = algorithmic flops (SURVEY 8(d): 80 per Student cell, exp = 30, log = 42) / time / measured DFMA "
                         "peak; it exceeds 1 because the kernel issues far fewer FP64 instructions per cell than that convention "
                         "charges.  cell_pipe_frac = cells really evaluated (strips shared between alphas count once) x ISSUED FP64 "
                         "instructions per cell (SASS count, tests/test_abi.py) / pipe slots in kernel_ms: a lower bound of ncu's sm__pipe_fp64_cycles_active (profiles/), which also "
                         "sees the axis stage, row set-up and masked lanes",
            "peak_source": f"measured in this run: dependency-free DFMA micro-benchmark, {peak_ms:.0f} ms "
                           "(MEASURED_PEAKS.json has no FP64 entry; nominal 148 SM x 64 lanes x 2 x 1.965 GHz = 37.2)",
        },
        "phases_us": {k: round(v, 2) for k, v in phases.items()},
        "pipelined": {"value": pipelined_value, "unit": UNIT,
                      "note": "ShardedSolver, two batches in flight: consecutive steps alternate between two streams (a plan "
                              "each), so a step's collective + finalize AND the partly idle end of its solve launch are covered "
                              "by the next step's solve; whole K-step region between two events, no L2 flush between steps"},
        "iterations": iters_np,
        "plan": {"ctas_per_sm": info.ctas_per_sm, "threads_per_cta": info.threads_per_cta,
                 "smem_bytes_per_cta": info.smem_bytes_per_cta, "sm_count": info.sm_count,
                 "kernel_variant": info.kernel_variant, "pow_octaves": info.pow_octaves, "chunk_days": int(info.chunk_days)},
    }
    if strong_rec:
        out["strong"] = strong_rec

    # ---- cpu_baseline (N = 1) + parity on a bounded sample spread over ALL shards (rank 0, any N) --------
    if args.cpu_sample_days > 0:
        from cvar_b200.backtest import exceedances      # the oracle itself is only executed inside the worker processes
        workers = len(os.sched_getaffinity(0))
        ndays = args.cpu_sample_days if world == 1 else min(args.cpu_sample_days, 64)
        sample = list(range(0, T_total, max(1, T_total // ndays)))[:ndays]
        pool = make_pool(workers)
        try:
            cpu_port_solve(inp_all, alphas[:1], sample[:workers], iters_np[:1], pool, workers)    # spin the workers up
            dt, cpu_var = cpu_port_solve(inp_all, alphas, sample, iters_np, pool, workers)
        finally:
            pool.shutdown()
        gpu_var = var_np if world > 1 else var_np            # (n_alpha, T_total) on every rank
        max_dvar = max(float(np.max(np.abs(gpu_var[k][sample] - cpu_var[a]))) for k, a in enumerate(alphas))
        rng = np.random.default_rng(11)
        r_ptf = rng.standard_normal(len(sample)) * 1.2
        exc_equal = all(exceedances(gpu_var[k][sample], r_ptf) == exceedances(cpu_var[a], r_ptf)
                        for k, a in enumerate(alphas))
        out["parity"] = {"max_abs_dvar_vs_oracle": max_dvar, "exceedance_counts_equal": bool(exc_equal),
                         "days_checked": len(sample), "shards_covered": world, "iterations": iters_np,
                         "note": "oracle re-solves the sampled days with the batch-wide iteration count the GPUs derived (Q7)"}
        if world == 1:
            out["cpu_baseline"] = {"value": len(sample) * na / dt, "unit": UNIT, "cores": workers, "kind": "port",
                                   "sample": f"{len(sample)} of {inp.T} days x {na} alpha(s), n={inp.n}, {dt:.1f} s wall"}
            if args.reference_unmodified:
                ref, ref_var = reference_unmodified_c1()
                if ref is None:
                    out["cpu_baseline"]["reference_unmodified"] = {"unavailable": ref_var}
                else:
                    # the same c1 batch on the GPU, for the parity of the two and the ratio at the reference's own size
                    from cvar_b200 import synthetic as syn
                    c1, c1_alphas = syn.baseline_config("c1")
                    with VarPlan(c1, device=local_rank) as c1_plan:
                        for _ in range(3):
                            c1_res = c1_plan.solve(c1.day_params(), c1_alphas, ptf_mean=c1.ptf_mean, details=False)
                        t0 = time.perf_counter()
                        for _ in range(20):
                            c1_plan.solve(c1.day_params(), c1_alphas, ptf_mean=c1.ptf_mean, details=False)
                        c1_s = (time.perf_counter() - t0) / 20
                    ref["b200_e2e_same_workload"] = {"value": c1.T / c1_s, "unit": UNIT}
                    ref["max_abs_dvar_b200_vs_reference"] = float(np.max(np.abs(c1_res.var[0] - ref_var)))
                    out["cpu_baseline"]["reference_unmodified"] = ref
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD)
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --days per GPU (default); strong: --days-total split over the GPUs")
    ap.add_argument("--days", type=int, default=None, help="days per GPU (default: the configuration's own T)")
    ap.add_argument("--days-total", type=int, default=None, help=f"total days with --scaling strong (default {STRONG_DAYS_TOTAL})")
    ap.add_argument("--n", type=int, default=None, help="grid points per axis (default: the configuration's own n)")
    ap.add_argument("--cpu-sample-days", type=int, default=128,
                    help="days of the workload the CPU port solves (cpu_baseline / parity legs, one --impl reference step)")
    ap.add_argument("--strong-steps", type=int, default=5, help="steps of the fixed-total measurement in the default line (0: off)")
    ap.add_argument("--no-reference-unmodified", dest="reference_unmodified", action="store_false",
                    help="skip timing the unmodified reference on c1 (N = 1, needs baseline/_ref)")
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
