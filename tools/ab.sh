#!/bin/bash
# In-box A/B of two builds of the library: tools/ab.sh <libA.so> <libB.so> [rounds] [extra gpu_quick_bench args]
# Alternates A and B so that clock / thermal drift hits both; prints the best step time and the per-launch table sums.
A=$1; B=$2; R=${3:-3}; shift 3 || true
mkdir -p gpurun_out
for r in $(seq 1 $R); do
  for v in A B; do
    L=$A; [ $v = B ] && L=$B
    CHB_LIB_PATH=$L timeout 300 python tools/gpu_quick_bench.py --steps 6 --warmup 2 --table "$@" > gpurun_out/ab_${v}_$r.log 2>&1
    echo "$v round $r: $(grep -E '^best' gpurun_out/ab_${v}_$r.log)  | $(grep -E 'per-launch' gpurun_out/ab_${v}_$r.log)"
  done
done
python - <<'PY'
import glob, re, collections
for v in "AB":
    acc = collections.defaultdict(list)
    for f in sorted(glob.glob("gpurun_out/ab_%s_*.log" % v)):
        for line in open(f):
            m = re.match(r"\s+(\S+)\s+([0-9.]+) ms", line)
            if m: acc[m.group(1)].append(float(m.group(2)))
    tot = sum(min(x) for x in acc.values())
    print(v, "sum of per-launch minima: %.3f ms" % tot)
    globals()["acc_" + v] = acc
print("launches where B differs from A by more than 3% (min over rounds):")
for k in acc_A:
    a, b = min(acc_A[k]), min(acc_B.get(k, [0]))
    if a > 0.05 and abs(b - a) / a > 0.03:
        print("  %-34s A %.3f  B %.3f  (%+.1f%%)" % (k, a, b, (b - a) / a * 100))
PY
