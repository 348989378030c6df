"""Record the DRAM traffic of one solve_kernel launch from an `ncu --set full` report into profiles/r2_traffic.json,
the file bench.py reads for `roofline.traffic` (an entry is only used while kernel variant and CTA size still match).

    python tools/ncu_traffic.py <workload> <report.ncu-rep> [<report> ...]
"""
import csv
import io
import json
import re
import subprocess
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
OUT = REPO / "profiles" / "r2_traffic.json"
workload, reports = sys.argv[1], sys.argv[2:]
table = json.loads(OUT.read_text()) if OUT.exists() else {}
for rep in reports:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}

    def nbytes(key):
        unit, val = d[key]
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
        return float(val.replace(",", "")) * scale

    name = d["Kernel Name"][1]
    variant = int(re.search(r"solve_kernel<(\d+)", name).group(1))
    head = subprocess.run(["git", "-C", str(REPO), "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    table[workload] = {
        "dram_bytes": int(nbytes("dram__bytes_read.sum") + nbytes("dram__bytes_write.sum")),
        "dram_bytes_read": int(nbytes("dram__bytes_read.sum")), "dram_bytes_write": int(nbytes("dram__bytes_write.sum")),
        "kernel_variant": variant, "threads_per_cta": int(float(d["launch__block_size"][1].replace(",", ""))),
        "grid": int(float(d["launch__grid_size"][1].replace(",", ""))),
        "source": f"profiles/{Path(rep).stem}_ncu_summary.txt (ncu --set full, one launch; captured at {head})",
    }
OUT.write_text(json.dumps(table, indent=1, sort_keys=True) + "\n")
print(json.dumps(table[workload]))
