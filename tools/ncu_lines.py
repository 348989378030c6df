"""Attribute ncu stall samples / executed instructions to CUDA source lines.

ncu's CSV source page is SASS-only; `nvdisasm -g` of the same cubin carries the line table.  Both list the
kernel's instructions in the same order, so they are joined by position.

    python tools/ncu_lines.py <report.ncu-rep> <lib.so> <kernel-substring> [top N]
"""
import collections
import csv
import io
import re
import subprocess
import sys
import tempfile
from pathlib import Path

rep, lib, pat = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
with tempfile.TemporaryDirectory() as tmp:
    subprocess.run(["cuobjdump", "-xelf", "all", str(Path(lib).resolve())], cwd=tmp, check=True, capture_output=True)
    cubin = next(Path(tmp).glob("*.cubin"))
    dis = subprocess.run(["nvdisasm", "-g", "-c", str(cubin)], capture_output=True, text=True).stdout
lines, inside, cur = [], False, ("?", 0)
for ln in dis.splitlines():
    if ln.startswith("//--------------------- .text."):
        inside = pat in ln
        continue
    if not inside:
        continue
    m = re.match(r'\s*//## File "(.*)", line (\d+)', ln)
    if m:
        cur = (Path(m.group(1)).name, int(m.group(2)))
        continue
    if re.match(r"\s*/\*[0-9a-f]{4,}\*/", ln):
        lines.append(cur)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
recs = []
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    try:
        recs.append((int(r[ix["Instructions Executed"]]), int(r[ix["# Samples"]]), r[ix["Source"]]))
    except ValueError:
        pass
if len(recs) != len(lines):
    print(f"warning: {len(recs)} profiled instructions vs {len(lines)} disassembled (library changed since the capture?)")
n = min(len(recs), len(lines))
agg = collections.defaultdict(lambda: [0, 0, 0])
ti = ts = 0
for (cnt, smp, txt), key in zip(recs[:n], lines[:n]):
    a = agg[key]
    a[0] += cnt
    a[1] += smp
    t = txt.split()
    op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    if op in ("DFMA", "DADD", "DMUL", "DSETP"):
        a[2] += cnt
    ti += cnt
    ts += smp
print(f"{'file:line':32s} {'instr %':>8s} {'samples %':>10s} {'fp64 share':>10s}")
for key, (cnt, smp, f64) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{key[0] + ':' + str(key[1]):32s} {100 * cnt / ti:8.2f} {100 * smp / ts:10.2f} {100 * f64 / max(cnt, 1):10.1f}")
