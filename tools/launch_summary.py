#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total, share."""
import collections, csv, re, sys

def main(path):
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except (KeyError, ValueError):
            continue
        u = row["Metric Unit"]
        v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v          # -> us
        name = re.sub(r"^void ", "", row["Kernel Name"])
        name = re.sub(r"\(.*", "", name)[:90]
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"{'total ms':>10} {'n':>6} {'share':>6} {'avg us':>9}  kernel")
    for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{v[1] / 1e3:10.3f} {v[0]:6d} {100 * v[1] / tot:5.1f}% {v[1] / v[0]:9.1f}  {k}")
    print(f"{tot / 1e3:10.3f} ms in {sum(v[0] for v in agg.values())} launches")

if __name__ == "__main__":
    main(sys.argv[1])
