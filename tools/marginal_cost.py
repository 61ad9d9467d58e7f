#!/usr/bin/env python
"""Marginal cost of each kernel family INSIDE the concurrent pipeline (32 streams of CUDA-graph replays).

Stand-alone kernel durations and instruction counts turned out to be poor predictors of the pipeline's throughput
(profiles/r2_sa_fused_limits.md), so this measures it directly: the C-ABI calls of one family are issued TWICE while
the graph is captured (every entry point is a pure function of its inputs, so the results do not change), and the
step time is compared with the unmodified pipeline.  The difference is what one more copy of that family costs per
step with everything else in flight -- i.e. what removing (or halving) it would buy.

    python tools/marginal_cost.py [--config 2] [--steps 100] [--families fps,sa_fused,...]
"""
import argparse
import json
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402

FAMILIES = {
    "fps": ("spc_furthest_point_sampling",),
    "sa_fused": ("spc_sa_fused_forward",),
    "pm_linear": ("spc_pm_linear",),
    "ball_query": ("spc_ball_query",),
    "three_nn_interp": ("spc_three_nn", "spc_interp_cat_pm", "spc_three_interpolate"),
}


def throughput(model, resident, steps, warmup, dup):
    from spacap3d_b200 import _lib
    from spacap3d_b200.pipeline import GraphedDetector
    orig = _lib.call
    count = [0]

    def call(name, *args):
        rc = orig(name, *args)
        if any(name.startswith(p) for p in dup):
            count[0] += 1
            orig(name, *args)
        return rc

    _lib.call = call
    try:
        runner = GraphedDetector(model, resident[0], n_streams=bench.N_STREAMS, result_keys=bench.RESULT_KEYS)
    finally:
        _lib.call = orig
    n_sets = len(resident)
    for i in range(warmup):
        runner.submit(resident[i % n_sets])
    runner.wait_all()
    cur = torch.cuda.current_stream()
    out = []
    for _ in range(9):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(cur)
        runner.fork_from(e0)
        for k in range(steps):
            runner.submit(resident[k % n_sets])
        runner.join_into(cur)
        e1.record(cur)
        torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1) / steps)
    runner.close()
    return statistics.median(out), count[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=2)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=40)
    ap.add_argument("--families", default=",".join(FAMILIES))
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    bench.CFG = dict(bench.CONFIGS[args.config], id=args.config)
    device = torch.device("cuda:0")
    torch.cuda.set_device(device)
    model = bench.make_detector(device)
    resident = [h.to(device) for h in bench.make_host_batches(0, 1)]
    spg = resident[0].shape[0]
    base, _ = throughput(model, resident, args.steps, args.warmup, ())
    rows = [{"family": "(none)", "ms_per_step": round(base, 4), "scenes_per_s": round(spg / base * 1e3, 1)}]
    print(json.dumps(rows[-1]), flush=True)
    for fam in args.families.split(","):
        ms, n = throughput(model, resident, args.steps, args.warmup, FAMILIES[fam])
        rows.append({"family": fam, "duplicated_calls_per_capture": n, "ms_per_step": round(ms, 4),
                     "marginal_us_per_step": round((ms - base) * 1e3, 1),
                     "share_of_step": round((ms - base) / base, 3)})
        print(json.dumps(rows[-1]), flush=True)
    base2, _ = throughput(model, resident, args.steps, args.warmup, ())
    rows.append({"family": "(none, repeated)", "ms_per_step": round(base2, 4)})
    print(json.dumps(rows[-1]), flush=True)
    if args.json:
        json.dump({"config": args.config, "rows": rows}, open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
