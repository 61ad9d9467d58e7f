#!/usr/bin/env python
"""Condense what tools/final_evidence.sh left in gpurun_out/ into the tracked summaries under profiles/ (round 2)."""
import csv
import glob
import gzip
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed"] + \
       ["smsp__average_warps_issue_stalled_%s_per_issue_active.ratio" % k for k in
        ("long_scoreboard", "short_scoreboard", "barrier", "wait", "branch_resolving", "math_pipe_throttle",
         "lg_throttle", "mio_throttle")]


def run(cmd, out=None):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if out:
        open(out, "w").write(r.stdout)
    return r.stdout


def raw_metrics(path):
    rows = list(csv.reader(open(path)))
    if len(rows) < 3:
        return []
    h, u = rows[0], rows[1]
    return [(row[h.index("Kernel Name")], {k: (row[h.index(k)], u[h.index(k)]) for k in KEYS if k in h}) for row in rows[2:]]


def main():
    py = sys.executable
    # model-level parity
    if os.path.exists(os.path.join(G, "parity.json")):
        shutil.copy(os.path.join(G, "parity.json"), os.path.join(P, "r2_reference_stack_parity.json"))
        run([py, os.path.join(ROOT, "tools", "parity_summary.py"), os.path.join(G, "parity.json")],
            os.path.join(P, "r2_reference_stack_parity.txt"))
    # bench lines
    os.makedirs(os.path.join(P, "r2_bench"), exist_ok=True)
    for f in sorted(glob.glob(os.path.join(G, "r2_bench_*.json")) + glob.glob(os.path.join(G, "r2_n*_config*.json"))):
        lines = open(f).read().strip().splitlines()
        if lines:
            open(os.path.join(P, "r2_bench", os.path.basename(f)), "w").write(lines[-1] + "\n")
    for name in ("r2_ops_vs_reference_graph.json", "r2_sanitizer_racecheck.log", "r2_sanitizer_memcheck.log",
                 "r2_train_step_profile.txt", "r2_sa_layers.json", "r2_marginal_cost_config2.json"):
        if os.path.exists(os.path.join(G, name)):
            shutil.copy(os.path.join(G, name), os.path.join(P, name))
    # ncu launch list of the bench command
    ll = os.path.join(G, "r2_launches_bench.csv")
    if os.path.exists(ll):
        with open(ll, "rb") as fi, gzip.open(os.path.join(P, "r2_launches_bench.csv.gz"), "wb") as fo:
            shutil.copyfileobj(fi, fo)
        run([py, os.path.join(ROOT, "tools", "launch_summary.py"), ll,
             "python bench.py --steps 2 --warmup 3 --no-cpu-baseline (first 3000 launches: eager warm-up + timed eager steps "
             "+ graph warm-up / capture)"], os.path.join(P, "r2_launches_bench_summary.csv"))
    # per-forward kernel table
    fw = os.path.join(G, "r2_forward.csv")
    if os.path.exists(fw):
        rows = list(csv.reader(l for l in open(fw) if l.startswith('"')))
        h = rows[0]
        names = {int(r[h.index("ID")]): r[h.index("Kernel Name")] for r in rows[1:]}
        ids = sorted(names)
        starts = [i for i in ids if "fps_morton_sort" in names[i]]
        n = len(ids) - starts[-1] + 2
        out = run([py, os.path.join(ROOT, "tools", "forward_table.py"), fw, str(n)])
        open(os.path.join(P, "r2_forward_table.csv"), "w").write("\n".join(l[:170] for l in out.splitlines()) + "\n")
    tr = os.path.join(G, "r2_sa_traffic.csv")
    if os.path.exists(tr):
        print(run([py, os.path.join(ROOT, "tools", "make_traffic_json.py"), tr, "sa_fused_pipe_kernel|sa_inline_kernel", "sa_fused",
                   "sa_common.cuh,sa_fused.cu,sa_inline.cu", "5"]).strip())
    # per-kernel ncu summaries
    out = ["# ncu --set full --clock-control none --import-source on, ONE launch per kernel (tools/profile_kernels.sh), B200.",
           "# group_points: the config-4 shape (B=8, C=132, 40 000 -> 2048 x 64; 727 MB algorithmic).  bqg_query: SA1 ball "
           "query (8 x 40 000, 2048 centres).",
           "# fps_*: 8 scenes x 40 000 -> 2048.  pm_linear: one FP / voting layer of the detector forward.  sa_fused_sa1 / "
           "_sa2: the SA1 (sa_inline_kernel: 64,64,128; nsample 64) and SA2 (sa_fused_pipe_kernel: 128,128,256; nsample 32) layers; "
           "group_grad_gather: the large-cloud backward of grouping at the config-4 shape.",
           "# three_interpolate: the FP2 shape.  Per-launch times are cold-cache and serialised."]
    met = {}
    for f in sorted(glob.glob(os.path.join(G, "r2_*.raw.csv"))):
        tag = os.path.basename(f).replace(".raw.csv", "")
        for name, m in raw_metrics(f):
            met[tag] = m
            out.append("=" * 100)
            out.append("%s   [%s]" % (name[:120], tag))
            out += ["  %-82s %s %s" % (k, v[0], v[1]) for k, v in m.items()]
    open(os.path.join(P, "r2_kernels_ncu.txt"), "w").write("\n".join(out) + "\n")
    c, b = met.get("r2_fps_cluster"), met.get("r2_fps_bucket")
    if c and b:
        txt = ["# SA1 sampler, 8 scenes x 40 000 points -> 2048 picks, ncu --set full (one launch each; cold, serialised), B200",
               "# before = round-1 design: fps_cluster_kernel<20,256> (everything on chip, 8-CTA clusters, DSMEM exchange)",
               "# after  = round-2 pipeline sampler: fps_bucket_kernel<64,16,1,2> (points parked in L2, one 512-thread CTA per scene)",
               "# (the round-1 pipeline ran the culled variant of the cluster kernel: 1.5 ms per call, 2.7 SMs per scene = 4.0 "
               "SM-ms per scene)", "", "%-64s %18s %18s" % ("metric", "cluster (before)", "bucketed (after)")]
        for k in KEYS:
            if k in c and k in b:
                short = k.replace("smsp__average_warps_issue_stalled_", "stall:").replace("_per_issue_active.ratio", "")
                txt.append("%-64s %18s %18s" % (short, c[k][0][:16], b[k][0][:16]))
        dc, db = float(c["gpu__time_duration.sum"][0]), float(b["gpu__time_duration.sum"][0])
        ic, ib = float(c["smsp__inst_executed.sum"][0]), float(b["smsp__inst_executed.sum"][0])
        txt += ["", "SM-time per scene (CTAs x duration / CTAs per SM / 8 scenes):",
                "  cluster : 64 CTAs x %.3f ms / 2 per SM / 8 = %.2f SM-ms per scene" % (dc, 64 * dc / 2 / 8),
                "  bucketed:  8 CTAs x %.3f ms / 2 per SM / 8 = %.2f SM-ms per scene" % (db, 8 * db / 2 / 8),
                "warp instructions per scene: cluster %.3g, bucketed %.3g" % (ic / 8, ib / 8),
                "sampling alone, CUDA-graph calls round-robin on 32 streams (tools/time_fps.py): cluster 18.4 k scenes/s, "
                "bucketed 56.6 k scenes/s",
                "detector pipeline (bench.py, config 2, 100 steps): 12.9 k scenes/s (round 1) -> 18.7 k scenes/s with this sampler "
                "(-> 20.5 k after round 2b's fused-SA / pm_linear changes); its marginal cost in the pipeline is 161-168 us of a "
                "389-412 us step (tools/marginal_cost.py) = the 8 CTAs x 2.95 ms it holds",
                "top stalls of the bucketed kernel: the per-round named barrier (39 %% of all stall samples sit right behind it: "
                "warps wait for the one with the most bucket visits), fixed-latency dependencies (wait), shared-memory "
                "latency (short scoreboard); issue slots %s %% busy." % b["smsp__issue_active.avg.pct_of_peak_sustained_active"][0][:5]]
        open(os.path.join(P, "r2_fps_ncu.txt"), "w").write("\n".join(txt) + "\n")
    print("profiles updated")


if __name__ == "__main__":
    main()
