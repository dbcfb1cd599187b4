#!/bin/bash
# In-box A/B of one environment switch of the library:  ROUNDS=3 tools/ab_env.sh CHB_PDL 0 1
VAR=$1; shift
R=${ROUNDS:-3}
mkdir -p gpurun_out
for r in $(seq 1 $R); do
  for v in "$@"; do
    env $VAR=$v timeout 300 python tools/gpu_quick_bench.py --steps 8 --warmup 3 > gpurun_out/abenv_${VAR}_${v}_$r.log 2>&1
    echo "$VAR=$v round $r: $(grep -E '^B=' gpurun_out/abenv_${VAR}_${v}_$r.log)"
  done
done
