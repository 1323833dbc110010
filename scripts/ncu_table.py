"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: per-kernel totals of the LAST of `runs` runs."""
import csv
import re
import sys


def main(path, runs=3, per_launch=True):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = [(x["Kernel Name"], float(x["Metric Value"].replace(",", "")), x.get("Grid Size")) for x in csv.DictReader(lines)]
    n = len(rows) // runs
    last = rows[-n:]
    tot = sum(v for _, v, _ in last)
    print("launches per run %d, total %.1f us" % (n, tot / 1000))
    agg = {}
    for k, v, g in last:
        k2 = re.sub(r"\(.*", "", k)
        a = agg.setdefault(k2, [0, 0])
        a[0] += v
        a[1] += 1
    for k, (v, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print("%-60s n=%3d %9.1f us %5.1f%%" % (k[:60], c, v / 1000, 100 * v / tot))
    if per_launch:
        for i, (k, v, g) in enumerate(last):
            print(i, re.sub(r"\(.*", "", k)[-40:], g, "%.1f us" % (v / 1000))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 3)
