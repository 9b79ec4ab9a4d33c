#!/usr/bin/env python
"""Compact per-kernel summary of an .ncu-rep (--set full): time, DRAM bytes, tensor/SM activity, occupancy,
top stall reasons. Usage: python tools/ncu_summary.py file.ncu-rep [--source N]"""
import csv, io, subprocess, sys

KEYS = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active%"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__occupancy_limit_registers", "occ_regs"), ("launch__occupancy_limit_shared_mem", "occ_smem"),
        ("launch__occupancy_limit_warps", "occ_warps"), ("sm__maximum_warps_per_active_cycle_pct", "theo_occ%"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_conflicts"),
        ("lts__t_sector_hit_rate.pct", "l2_hit%"), ("sm__cycles_elapsed.max", "cycles"),
        ("smsp__inst_executed.sum", "warp_insts")]


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def main():
    path = sys.argv[1]
    hdr, units, rows = raw(path)
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows:
        print("==", r[idx["Kernel Name"]][:110])
        parts = []
        for k, nm in KEYS:
            if k in idx:
                parts.append(f"{nm}={r[idx[k]]}{units[idx[k]] if units[idx[k]] not in ('', 'register/thread', 'block') else ''}")
        print("   " + "  ".join(parts))
        stalls = []
        for h, i in idx.items():
            if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued"):
                try:
                    stalls.append((int(r[i]), h.replace("smsp__pcsamp_warps_issue_stalled_", "")))
                except ValueError:
                    pass
        tot = sum(s for s, _ in stalls) or 1
        stalls.sort(reverse=True)
        print("   stalls: " + "  ".join(f"{nm}={100 * s / tot:.0f}%" for s, nm in stalls[:7]))
    if "--source" in sys.argv:
        n = int(sys.argv[sys.argv.index("--source") + 1])
        out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        # find header
        for hi, r in enumerate(rows):
            if "Source" in r and any("Sampling" in c for c in r):
                break
        hdr = rows[hi]
        si = hdr.index("Source")
        ci = [i for i, c in enumerate(hdr) if c.startswith("# Samples") or c == "Warp Stall Sampling (All Samples)" or c == "Warp Stall Sampling (All Cycles)"]
        ci = ci[0] if ci else None
        data = []
        for r in rows[hi + 1:]:
            try:
                data.append((int(r[ci]), r[si]))
            except (ValueError, IndexError, TypeError):
                pass
        tot = sum(d[0] for d in data) or 1
        data.sort(reverse=True)
        for s, src in data[:n]:
            print(f"   {100 * s / tot:5.1f}%  {src[:120]}")


if __name__ == "__main__":
    main()
