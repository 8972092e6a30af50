#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel name the
launch count, total and mean device time and the share of the whole list.  Optionally restrict
to the launches of ONE steady-state step (the last N launches, N = launches per step).

    python tools/summarize_launches.py gpurun_out/launches.csv [--last N] > profiles/....md
"""
import csv
import re
import sys
from collections import OrderedDict


def main():
    path = sys.argv[1]
    last = int(sys.argv[sys.argv.index("--last") + 1]) if "--last" in sys.argv else 0
    rows = []
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*$", "", r["Kernel Name"]).replace("void ", "").replace("pp::", "")
        rows.append((name, float(r["Metric Value"]) / 1e3, r["Grid Size"], r["Block Size"]))
    if last:
        rows = rows[-last:]
    agg = OrderedDict()
    for name, us, grid, block in rows:
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += us
    total = sum(a[1] for a in agg.values())
    print(f"source: {path}; launches: {len(rows)}" + (f" (last {last} = one steady-state step)" if last else "") +
          f"; total device time {total:.1f} us (ncu per-launch times are cold-cache and serialised: compare shares)")
    print()
    print("| kernel | launches | total us | mean us | share |")
    print("|---|---:|---:|---:|---:|")
    for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{name}` | {n} | {us:.1f} | {us / n:.1f} | {us / total:.1%} |")


if __name__ == "__main__":
    main()
