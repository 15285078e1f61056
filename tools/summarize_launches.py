#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name.
usage: python tools/summarize_launches.py gpurun_out/launches.csv > profiles/launches_rNN.md"""
import collections
import csv
import re
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1000.0 if unit.startswith("n") else (v * 1000.0 if unit.startswith("m") else v)
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"# ncu launch list summary of `{path}` (cold-cache, serialised: compare SHARES)\n")
    print(f"total device time {tot:.1f} us over {sum(v[0] for v in agg.values())} launches\n")
    print("| share | total us | launches | us/launch | kernel |")
    print("|---:|---:|---:|---:|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {100 * v[1] / tot:.1f}% | {v[1]:.1f} | {v[0]} | {v[1] / v[0]:.2f} | `{k[:100]}` |")


if __name__ == "__main__":
    main(sys.argv[1])
