"""Per-launch table of an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` log:
python tools/launch_table.py <csv> [first_id last_id]   -> aggregate by kernel name, then the launch sequence."""
import collections
import csv
import sys


def load(path):
    lines = [ln for ln in open(path) if ln.startswith('"')]
    rows = collections.OrderedDict()
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for r in csv.DictReader(lines):
        d = rows.setdefault(int(r["ID"]), {"name": r["Kernel Name"], "grid": r["Grid Size"], "ms": 0.0, "rd": 0.0, "wr": 0.0})
        v = float(r["Metric Value"].replace(",", ""))
        m, u = r["Metric Name"], r["Metric Unit"]
        if m.startswith("gpu__time"):
            d["ms"] = v / 1e6 if u in ("ns", "nsecond") else (v / 1e3 if u in ("us", "usecond") else v)
        elif "read" in m:
            d["rd"] = v * scale.get(u, 1) / 1e6
        else:
            d["wr"] = v * scale.get(u, 1) / 1e6
    return rows


def short(n):
    return n.replace("void ", "").replace("chb::", "")[:58]


def main():
    rows = load(sys.argv[1])
    ids = list(rows)
    if len(sys.argv) > 3:
        ids = [i for i in ids if int(sys.argv[2]) <= i <= int(sys.argv[3])]
    agg = collections.OrderedDict()
    for i in ids:
        a = agg.setdefault(short(rows[i]["name"]), [0, 0.0, 0.0, 0.0])
        a[0] += 1; a[1] += rows[i]["ms"]; a[2] += rows[i]["rd"]; a[3] += rows[i]["wr"]
    tot = sum(a[1] for a in agg.values())
    print("%d launches, %.3f ms (ncu: serialised, cold cache)" % (len(ids), tot))
    for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        gbs = (a[2] + a[3]) / 1e3 / (a[1] * 1e-3) if a[1] > 0 else 0.0
        print("%-60s n=%3d %8.3f ms %5.1f%%  rd %8.1f MB  wr %8.1f MB  %6.0f GB/s" % (n, a[0], a[1], 100 * a[1] / tot, a[2], a[3], gbs))
    print()
    for i in ids:
        r = rows[i]
        print("%4d %-58s %8.4f ms rd %8.1f wr %8.1f grid %s" % (i, short(r["name"]), r["ms"], r["rd"], r["wr"], r["grid"]))


if __name__ == "__main__":
    main()
