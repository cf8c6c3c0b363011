#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: share of device time per kernel."""
import collections
import csv
import re
import sys


def main(path, top=18):
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0, []])
    tot = 0.0
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1e3 if unit == "ns" else (v * 1e3 if unit == "ms" else v)
        name = re.sub(r"\(.*", "", row["Kernel Name"])[:70]
        agg[name][0] += 1
        agg[name][1] += v
        agg[name][2].append(round(v))
        tot += v
    print(f"total {tot:.0f} us over {sum(a[0] for a in agg.values())} launches")
    for k, (c, t, l) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{t:10.1f} us {100 * t / tot:5.1f}% n={c:4d} {k}  max {sorted(l)[-3:]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 18)
