#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (time share of the step)."""
import collections
import csv
import sys


def main(path):
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except (ValueError, KeyError):
            continue
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(row["Metric Unit"], 1e-6)
        k = row["Kernel Name"].split("(")[0][:60]
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"# {path}: {sum(v[0] for v in agg.values())} launches, {tot:.2f} ms total (cold-cache, serialised: compare shares)")
    print(f"{'ms':>10} {'share':>7} {'launches':>8}  kernel")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{v[1]:10.3f} {100 * v[1] / tot:6.2f}% {v[0]:8d}  {k}")


if __name__ == "__main__":
    main(sys.argv[1])
