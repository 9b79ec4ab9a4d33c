#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv` launch
list: per-kernel count, total time, share of the step, DRAM bytes per launch.
    python tools/launch_summary.py launches.csv [--json out.json]"""
import collections, csv, json, re, sys

UNIT = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main(path, json_out=None):
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    per = collections.OrderedDict()          # launch id -> dict
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except (KeyError, ValueError):
            continue
        v *= UNIT.get(row["Metric Unit"], 1.0)
        name = re.sub(r"^void ", "", row["Kernel Name"])
        name = re.sub(r"\(.*", "", name)[:90]
        d = per.setdefault(row["ID"], {"name": name})
        d[row["Metric Name"]] = v
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
    for d in per.values():
        a = agg[d["name"]]
        a[0] += 1
        a[1] += d.get("gpu__time_duration.sum", 0.0)
        a[2] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
    tot = sum(v[1] for v in agg.values())
    print(f"{'total ms':>10} {'n':>6} {'share':>6} {'avg us':>9} {'DRAM MB/launch':>15}  kernel")
    for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{v[1] / 1e3:10.3f} {v[0]:6d} {100 * v[1] / tot:5.1f}% {v[1] / v[0]:9.1f} {v[2] / v[0] / 1e6:15.1f}  {k}")
    print(f"{tot / 1e3:10.3f} ms in {sum(v[0] for v in agg.values())} launches (cold-cache, serialised: compare shares)")
    gem = [v for k, v in agg.items() if k.startswith("fvqa::gemm_nt_pair_kernel") or k.startswith("gemm_nt_pair_kernel")]
    if gem and json_out:
        n = sum(v[0] for v in gem)
        out = {"avg_dram_bytes_per_launch": sum(v[2] for v in gem) / n, "launches": n,
               "avg_us_per_launch_under_ncu": sum(v[1] for v in gem) / n, "share_of_step_under_ncu": sum(v[1] for v in gem) / tot,
               "note": "dram__bytes_read.sum + dram__bytes_write.sum averaged over the gemm_nt_pair_kernel launches of one 7B NExT-QA "
                       "step (ncu launch list, cold cache per launch)", "source": path}
        with open(json_out, "w") as f:
            json.dump(out, f, indent=1)
        print("wrote", json_out, out)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[sys.argv.index("--json") + 1] if "--json" in sys.argv else None)
