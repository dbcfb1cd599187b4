#!/bin/bash
# tools/exp_env.sh "<grep pattern>" "VAR=..." "VAR=..." ...: per-launch table under different environment settings, two rounds, alternating order
PAT=$1; shift
for r in 1 2; do
  for e in "$@"; do
    echo "== round $r: $e"
    env $e timeout 300 python tools/gpu_quick_bench.py --steps 5 --warmup 2 --table 2>&1 | grep -E "best|total|$PAT"
  done
done
