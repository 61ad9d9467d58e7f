#!/usr/bin/env python
"""Single-call latency and multi-stream throughput of the FPS samplers (8 scenes x 40 000 -> 2048 by default).

    python tools/time_fps.py [--points 40000] [--npoint 2048] [--batch 8] [--streams 1,4,16]

Throughput mode submits `--batch`-scene calls round-robin on S streams (what the graph pipeline does) and reports
calls/ms; since the samplers are latency-bound, calls/ms at S streams divided by S shows how much of the GPU one
call holds."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", type=int, default=40000)
    ap.add_argument("--npoint", type=int, default=2048)
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--streams", default="1,4,16,32")
    ap.add_argument("--algos", default="cluster,bucket")
    args = ap.parse_args()
    from spacap3d_b200 import _ext
    from spacap3d_b200.scenes import make_scene_xyz
    dev = torch.device("cuda:0")
    xyz = torch.from_numpy(np.stack([make_scene_xyz(2000 + i, args.points) for i in range(args.batch)], 0)).to(dev)
    want = None
    for algo in args.algos.split(","):
        code = {"auto": _ext.FPS_AUTO, "cluster": _ext.FPS_CLUSTER, "bucket": _ext.FPS_BUCKET}[algo]
        with _ext.launch_options(fps_algo=code):
            for _ in range(3):
                idx, _ = _ext.furthest_point_sampling_with_xyz(xyz, args.npoint)
            torch.cuda.synchronize()
            if want is None:
                want = idx.clone()
            same = bool(torch.equal(idx, want))
            ts = []
            for _ in range(10):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                _ext.furthest_point_sampling_with_xyz(xyz, args.npoint)
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            rec = {"algo": algo, "same_indices": same, "single_call_ms": round(float(np.median(ts)), 4),
                   "us_per_round": round(float(np.median(ts)) * 1e3 / (args.npoint - 1), 4)}
            for S in [int(x) for x in args.streams.split(",")]:
                streams = [torch.cuda.Stream() for _ in range(S)]
                graphs = []
                for st in streams:                      # one captured call per stream (no launch overhead in the loop)
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, stream=st):
                        _ext.furthest_point_sampling_with_xyz(xyz, args.npoint)
                    graphs.append(g)
                torch.cuda.synchronize()
                reps = 6
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                cur = torch.cuda.current_stream()
                e0.record(cur)
                for st in streams:
                    st.wait_event(e0)
                for r in range(reps):
                    for st, g in zip(streams, graphs):
                        with torch.cuda.stream(st):
                            g.replay()
                for st in streams:
                    ev = torch.cuda.Event()
                    ev.record(st)
                    cur.wait_event(ev)
                e1.record(cur)
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1)
                rec["streams_%d_calls_per_ms" % S] = round(reps * S / ms, 3)
                rec["streams_%d_scenes_per_s" % S] = round(reps * S * args.batch / ms * 1e3, 1)
                del graphs
            print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
