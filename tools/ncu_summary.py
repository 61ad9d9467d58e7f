#!/usr/bin/env python
"""Summarise an .ncu-rep: per-kernel key metrics + top stall lines (needs -lineinfo + --import-source).
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--top 12]
"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]


def run(args):
    return subprocess.run(["ncu", "-i"] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 12
    raw = list(csv.reader(io.StringIO(run([rep, "--page", "raw", "--csv"]))))
    hdr, units = raw[0], raw[1]
    for row in raw[2:]:
        print("=" * 100)
        print(row[hdr.index("Kernel Name")][:110])
        for k in KEYS:
            if k in hdr:
                print("  %-70s %s %s" % (k, row[hdr.index(k)], units[hdr.index(k)]))
    src = list(csv.reader(io.StringIO(run([rep, "--page", "source", "--csv"]))))
    kern, cur = [], None
    for r in src:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            kern.append(cur)
        elif r and r[0] == "Address":
            cur["hdr"] = r
        elif cur is not None and r:
            cur["rows"].append(r)
    seen = set()
    for k in kern:
        if k["name"] in seen or "hdr" not in k:
            continue
        seen.add(k["name"])
        h = k["hdr"]
        si, ci, ei = h.index("Warp Stall Sampling (All Samples)"), h.index("Source"), h.index("Instructions Executed")
        tot = sum(float(r[si] or 0) for r in k["rows"]) or 1.0
        print("-" * 100)
        print("stall samples:", k["name"][:90], int(tot))
        rows = sorted(enumerate(k["rows"]), key=lambda t: -float(t[1][si] or 0))[:top]
        for i, r in sorted(rows):
            print("  %5d %5.1f%% exec=%-9s %s" % (i, 100 * float(r[si] or 0) / tot, r[ei], r[ci].strip()[:90]))


if __name__ == "__main__":
    main()
