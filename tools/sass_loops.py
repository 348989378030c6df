"""List the loops of a kernel in `cuobjdump -sass` output with their instruction mix (FP64 vs the rest).

    cuobjdump -sass lib.so | python tools/sass_loops.py solve_kernelILi1
"""
import re
import sys
from collections import Counter

pat = sys.argv[1]
lines = sys.stdin.read().splitlines()
inside = False
ins = []
for ln in lines:
    if "Function :" in ln:
        inside = pat in ln
        continue
    if not inside:
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        ins.append((int(m.group(1), 16), m.group(2)))
addr_index = {a: i for i, (a, _) in enumerate(ins)}
loops = []
for i, (a, txt) in enumerate(ins):
    m = re.search(r"BRA(?:\.\w+)*\s+(?:`\(\.L_x_\d+\)|0x([0-9a-f]+))", txt)
    if m and m.group(1):
        tgt = int(m.group(1), 16)
        if tgt <= a and tgt in addr_index:
            loops.append((addr_index[tgt], i))
FP64 = ("DFMA", "DADD", "DMUL", "DSETP", "DMNMX")
order = sorted(loops, key=lambda t: t[0] - t[1]) if len(sys.argv) < 3 else sorted(loops, key=lambda t: t[1] - t[0])
for s, e in order[:int(sys.argv[2]) if len(sys.argv) > 2 else 12]:
    ops = Counter()
    for _, txt in ins[s:e + 1]:
        t = txt.split()
        op = t[1] if t[0].startswith("@") else t[0]
        ops[op.split(".")[0]] += 1
    n = e - s + 1
    f = sum(ops[o] for o in FP64)
    print(f"loop {ins[s][0]:#06x}..{ins[e][0]:#06x}  {n:4d} instr, FP64 {f:4d}, other {n - f:4d} :: " +
          ", ".join(f"{k}={v}" for k, v in ops.most_common(14)))
