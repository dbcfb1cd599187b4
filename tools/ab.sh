#!/bin/bash
# In-box A/B/... of several builds of the library:  ROUNDS=3 tools/ab.sh <lib1.so> <lib2.so> ...
# Alternates the builds so that clock / thermal drift hits all of them; prints best step times and per-launch minima.
R=${ROUNDS:-3}
mkdir -p gpurun_out; rm -f gpurun_out/ab_*.log
for r in $(seq 1 $R); do
  n=$#
  for j in $(seq 0 $((n-1))); do
    i=$j; [ $((r % 2)) -eq 0 ] && i=$((n-1-j))   # even rounds run the builds in reverse order
    L=${@:$((i+1)):1}
    CHB_LIB_PATH=$L timeout 300 python tools/gpu_quick_bench.py --steps 6 --warmup 2 --table > gpurun_out/ab_${i}_$r.log 2>&1
    echo "[$i] round $r: $(grep -E '^best' gpurun_out/ab_${i}_$r.log | cut -c1-40)  | $(grep -E 'per-launch' gpurun_out/ab_${i}_$r.log | sed 's/.*total/total/')"
  done
done
python - "$@" <<'PY'
import glob, re, collections, sys
libs = sys.argv[1:]
accs = []
for i, lib in enumerate(libs):
    acc = collections.OrderedDict()
    for f in sorted(glob.glob("gpurun_out/ab_%d_*.log" % i)):
        for line in open(f):
            m = re.match(r"\s+(\S+)\s+([0-9.]+) ms", line)
            if m: acc.setdefault(m.group(1), []).append(float(m.group(2)))
    accs.append(acc)
    print("[%d] %-40s sum of per-launch minima: %.3f ms" % (i, lib.split('/')[-1], sum(min(x) for x in acc.values())))
print("per-launch minima (ms) of launches > 0.3 ms:")
for k in accs[0]:
    if min(accs[0][k]) > 0.3:
        print("  %-34s %s" % (k, "  ".join("%.3f" % min(a.get(k, [0])) for a in accs)))
PY
