"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals/shares."""
import collections
import csv
import sys


def load(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    seq = []
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1000 if unit == "ns" else (v * 1000 if unit == "ms" else v)
        seq.append((row["Kernel Name"], row["Grid Size"], v))
    return seq


def main():
    seq = load(sys.argv[1])
    agg = collections.OrderedDict()
    for name, grid, v in seq:
        a = agg.setdefault(name[:70], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print("%-72s %5s %10s %9s %6s" % ("kernel", "n", "total_us", "avg_us", "share"))
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-72s %5d %10.1f %9.1f %6.3f" % (k, n, t, t / n, t / tot))
    print("total_us %.1f over %d launches" % (tot, len(seq)))
    if len(sys.argv) > 2:
        for name, grid, v in seq:
            if sys.argv[2] in name:
                print(name[:60], grid, "%.1f" % v)


if __name__ == "__main__":
    main()
