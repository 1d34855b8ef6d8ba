#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (share of the step)."""
import collections
import csv
import sys


def main():
    lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
    rows = list(csv.DictReader(lines))
    agg = collections.OrderedDict()
    for r in rows:
        try:
            v = float(r["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        k = r["Kernel Name"].split("(")[0]
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values()) or 1.0
    print(f"total {tot / 1e6:.3f} ms over {sum(a[0] for a in agg.values())} launches")
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{t / 1e6:10.3f} ms {n:4d} launches {100 * t / tot:5.1f}%  {k[:100]}")


if __name__ == "__main__":
    main()
