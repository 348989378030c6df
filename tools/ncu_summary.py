"""Summarise an .ncu-rep (first kernel): key raw metrics + SASS opcode mix + per-opcode stall samples.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [out.txt]
"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
out = open(sys.argv[2], "w") if len(sys.argv) > 2 else sys.stdout
KEEP = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__cycles_active.avg", "sm__cycles_elapsed.avg", "smsp__inst_executed.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
d = dict(zip(hdr, zip(units, vals)))
print(f"report: {rep}\nkernel: {d.get('Kernel Name', ('', '?'))[1]}\n", file=out)
for k in KEEP:
    if k in d:
        print(f"{k} [{d[k][0]}] = {d[k][1]}", file=out)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
byop, samp = collections.Counter(), collections.Counter()
tot = ts = 0
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    try:
        n, s = int(r[ix["Instructions Executed"]]), int(r[ix["# Samples"]])
    except ValueError:
        continue
    t = r[ix["Source"]].split()
    op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    byop[op] += n
    samp[op] += s
    tot += n
    ts += s
print(f"\nSASS opcode mix (warp instructions executed, total {tot}; stall samples {ts})", file=out)
fp64 = sum(byop[o] for o in ("DFMA", "DADD", "DMUL", "DSETP", "DMNMX"))
print(f"FP64-pipe instructions: {fp64} = {100 * fp64 / tot:.1f} % of issued", file=out)
for op, n in byop.most_common(22):
    print(f"  {op:10s} {n:13d} {100 * n / tot:5.1f} %   samples {100 * samp[op] / max(ts, 1):5.1f} %", file=out)
