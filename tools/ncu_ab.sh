#!/bin/bash
# ncu --set full capture of the solve kernel for several builds:  tools/ncu_ab.sh <workload> <tag-prefix> lib1.so lib2.so ...
# (run under gpurun; the plain command runs first, as the profiling recipe requires)
w="$1"; pre="$2"; shift; shift
for lib in "$@"; do
  tag=$(basename $lib .so)
  export CVAR_B200_LIB=$lib
  python bench.py --workload $w --steps 2 --warmup 3 --cpu-sample-days 0 > gpurun_out/plain_${tag}.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:solve_kernel -s 2 -c 1 -f -o gpurun_out/${pre}_${w}_${tag} \
      python bench.py --workload $w --steps 2 --warmup 3 --cpu-sample-days 0 > gpurun_out/ncu_${tag}.log 2>&1
  echo "$w $tag rc=$?"
done
