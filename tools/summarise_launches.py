"""Summarises an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` launch list.

python tools/summarise_launches.py gpurun_out/<tag>_launches_time_dram.csv profiles/<tag>   [--launches-per-step 57]

Writes <out>_launches_time_dram.csv (a copy), <out>_launch_summary.txt (per kernel name: launches, time share, DRAM
bytes) and refreshes profiles/traffic.json (DRAM bytes per conv launch, averaged over one step) for bench.py.
"""
import collections
import csv
import json
import os
import re
import shutil
import sys


def main():
    src, out = sys.argv[1], sys.argv[2]
    per_step = 57
    if "--launches-per-step" in sys.argv:
        per_step = int(sys.argv[sys.argv.index("--launches-per-step") + 1])
    rows = collections.OrderedDict()
    with open(src) as f:
        lines = [ln for ln in f if ln.startswith('"')]
    for r in csv.DictReader(lines):
        d = rows.setdefault(int(r["ID"]), {"name": r["Kernel Name"], "grid": r["Grid Size"], "block": r["Block Size"]})
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        m = r["Metric Name"]
        if m.startswith("dram__bytes"):
            v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
        elif m.startswith("gpu__time"):
            v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)  # -> ms
        d[m] = v
    ids = sorted(rows)
    short = lambda n: re.sub(r"\(.*", "", n)  # noqa: E731
    agg = collections.OrderedDict()
    for i in ids:
        d = rows[i]
        a = agg.setdefault(short(d["name"]), {"n": 0, "ms": 0.0, "rd": 0.0, "wr": 0.0})
        a["n"] += 1
        a["ms"] += d.get("gpu__time_duration.sum", 0.0)
        a["rd"] += d.get("dram__bytes_read.sum", 0.0)
        a["wr"] += d.get("dram__bytes_write.sum", 0.0)
    tot = sum(a["ms"] for a in agg.values())
    txt = ["source: %s (%d launches captured; ncu per-launch times are serialised and cold-cache: compare SHARES)" %
           (os.path.basename(src), len(ids)), "",
           "%-64s %6s %10s %7s %12s %12s" % ("kernel", "n", "ms", "share", "dram rd GB", "dram wr GB")]
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
        txt.append("%-64s %6d %10.3f %6.1f%% %12.3f %12.3f" % (k[:64], a["n"], a["ms"], 100 * a["ms"] / tot,
                                                              a["rd"] / 1e9, a["wr"] / 1e9))
    # one step = the last `per_step` conv launches in the capture
    conv = [i for i in ids if "conv_igemm_kernel" in rows[i]["name"]]
    if len(conv) >= per_step:
        last = conv[-per_step:]
        ms = sum(rows[i].get("gpu__time_duration.sum", 0.0) for i in last)
        by = sum(rows[i].get("dram__bytes_read.sum", 0.0) + rows[i].get("dram__bytes_write.sum", 0.0) for i in last)
        lo, hi = last[0], last[-1]
        other = sum(rows[i].get("gpu__time_duration.sum", 0.0) for i in ids if lo <= i <= hi and i not in set(last))
        txt += ["", "last full step: %d conv launches, %.3f ms (ncu serialised), %.2f GB DRAM traffic; other kernels inside "
                "the step %.3f ms (%.2f%%)" % (per_step, ms, by / 1e9, other, 100 * other / (ms + other))]
        txt.append("top conv launches of that step:")
        for i in sorted(last, key=lambda i: -rows[i].get("gpu__time_duration.sum", 0.0))[:12]:
            d = rows[i]
            txt.append("  id %4d grid %-14s %8.3f ms  rd %7.1f MB  wr %7.1f MB" %
                       (i, d["grid"], d.get("gpu__time_duration.sum", 0.0), d.get("dram__bytes_read.sum", 0.0) / 1e6,
                        d.get("dram__bytes_write.sum", 0.0) / 1e6))
        tj = {"dram_bytes_per_step": by, "conv_launches_per_step": per_step, "dram_bytes_per_launch_avg": by / per_step,
              "source": "profiles/%s_launches_time_dram.csv (ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,"
                        "dram__bytes_write.sum --clock-control none, B=64 256x256, last full step of the capture)" %
                        os.path.basename(out)}
        with open(os.path.join(os.path.dirname(out) or ".", "traffic.json"), "w") as f:
            json.dump(tj, f, indent=1)
    shutil.copyfile(src, out + "_launches_time_dram.csv")
    with open(out + "_launch_summary.txt", "w") as f:
        f.write("\n".join(txt) + "\n")
    print("\n".join(txt))


if __name__ == "__main__":
    main()
