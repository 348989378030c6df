#!/bin/bash
# A/B bench of several builds of libcvar_b200.so on one GPU box:  tools/ab_bench.sh "<workloads>" lib1.so lib2.so ...
# Prints workload, library, ms/step, solves/s and the parity record of each run (bench.py lines go to gpurun_out/ab/).
wl="$1"; shift
mkdir -p gpurun_out/ab
for w in $wl; do
  for lib in "$@"; do
    tag=$(basename $lib .so)
    CVAR_B200_LIB=$lib python bench.py --workload $w --steps 10 --warmup 3 --cpu-sample-days ${CPU_DAYS:-16} > gpurun_out/ab/${w}_${tag}.json 2> gpurun_out/ab/${w}_${tag}.err
    python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/ab/${w}_${tag}.json").read().strip().splitlines()[-1])
    print("${w}", "${tag}", "ms/step %.4f" % d["ms_per_step"], "kernel_ms %.4f" % d["roofline"]["kernel_ms"], "value %.4g" % d["value"], "e2e %.4g" % d["e2e"]["value"], d.get("parity"))
except Exception as ex:
    print("${w}", "${tag}", "FAILED", ex); print(open("gpurun_out/ab/${w}_${tag}.err").read()[-2000:])
PY
  done
done
