#!/usr/bin/env python
"""Per-op timing on BASELINE shapes: our sm_100a kernels vs the reference's own CUDA ops
(oracle/_ref, when present).  CUDA events, L2 flushed between iterations, median of N.

    python tools/bench_ops.py [--ops fps,ball_query,...] [--batch 8] [--iters 20] [--json out.json]
"""
import argparse
import json
import os
import statistics
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from spacap3d_b200 import _ext  # noqa: E402
from spacap3d_b200.scenes import make_scene_xyz  # noqa: E402


def timeit(fn, iters, flush):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        flush()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return statistics.median(ts), min(ts)


def timeit_graph(fn, iters, flush, reps=8):
    """Launch-overhead-free timing for ops that take a few microseconds: [flush, fn] x reps captured in one CUDA
    graph, minus the same graph without fn.  Returns (per-call ms, per-call ms) like timeit()."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    side = torch.cuda.Stream()

    def capture(with_fn):
        g = torch.cuda.CUDAGraph()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.graph(g, stream=side):
            for _ in range(reps):
                flush()
                if with_fn:
                    fn()
        return g

    ga, gb = capture(True), capture(False)

    def run(g):
        ts = []
        for _ in range(max(3, iters // 2)):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            g.replay()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return statistics.median(ts)

    run(ga), run(gb)
    d = max(run(ga) - run(gb), 0.0) / reps
    return d, d


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--graph", action="store_true",
                    help="time OUR ops inside a CUDA graph (no host launch gap); the reference ext stays eager")
    ap.add_argument("--ops", default="fps,ball_query,group,interp,three_nn")
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--iters", type=int, default=15)
    ap.add_argument("--points", type=int, default=40000)
    ap.add_argument("--json", default=None)
    ap.add_argument("--no-ref", action="store_true")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    timeit_ours = timeit_graph if args.graph else timeit
    ref = None
    if not args.no_ref:
        try:
            from oracle.build_ref import load_ref
            ref = load_ref()
        except Exception as e:  # noqa: BLE001
            print("reference ext unavailable:", e)
    B, N = args.batch, args.points
    fl = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    flush = lambda: fl.fill_(0)
    xyz = torch.from_numpy(np.stack([make_scene_xyz(50 + i, N) for i in range(B)], 0)).to(dev)
    out = []

    def rec(op, shape, ours, theirs, algo_bytes=None, extra=None):
        r = {"op": op, "shape": shape, "ours_ms": round(ours[0], 4), "ours_min_ms": round(ours[1], 4)}
        if theirs:
            r["ref_ms"] = round(theirs[0], 4)
            r["speedup"] = round(theirs[0] / ours[0], 2)
        if algo_bytes:
            r["algo_gbs"] = round(algo_bytes / ours[0] / 1e6, 1)
        if extra:
            r.update(extra)
        out.append(r)
        print(json.dumps(r), flush=True)

    ops = args.ops.split(",")
    sa = [(N, 2048, 0.2, 64), (2048, 1024, 0.4, 32), (1024, 512, 0.8, 16), (512, 256, 1.2, 16)]
    cur = xyz
    levels = []
    for (n, npoint, r, ns) in sa:
        idx, new_xyz = _ext.furthest_point_sampling_with_xyz(cur, npoint)
        levels.append((cur, new_xyz, npoint, r, ns))
        cur = new_xyz
    if "fps" in ops:
        for (pts, _, npoint, _, _) in levels:
            n = pts.shape[1]
            o = timeit_ours(lambda: _ext.furthest_point_sampling(pts, npoint), args.iters, flush)
            t = timeit(lambda: ref.furthest_point_sampling(pts, npoint), max(3, args.iters // 3), flush) if ref else None
            rec("fps", [B, n, npoint], o, t, extra={"us_per_round": round(o[0] * 1e3 / (npoint - 1), 4)})
    if "ball_query" in ops:
        for (pts, new_xyz, npoint, r, ns) in levels:
            n = pts.shape[1]
            o = timeit_ours(lambda: _ext.ball_query(new_xyz, pts, r, ns), args.iters, flush)
            t = timeit(lambda: ref.ball_query(new_xyz, pts, r, ns), max(3, args.iters // 3), flush) if ref else None
            rec("ball_query", [B, n, npoint, r, ns], o, t, B * (12 * n + 12 * npoint + 4 * npoint * ns))
    if "group" in ops:
        for (C, li) in ((3, 0), (1, 0), (7, 0), (132, 0), (128, 1), (256, 2), (256, 3)):
            pts, new_xyz, npoint, r, ns = levels[li]
            n = pts.shape[1]
            idx = _ext.ball_query(new_xyz, pts, r, ns)
            feats = torch.randn(B, C, n, device=dev)
            o = timeit_ours(lambda: _ext.group_points(feats, idx), args.iters, flush)
            t = timeit(lambda: ref.group_points(feats, idx), max(3, args.iters // 3), flush) if ref else None
            rec("group_points", [B, C, n, npoint, ns], o, t, 4 * B * (C * n + npoint * ns + C * npoint * ns))
            g = torch.randn(B, C, npoint, ns, device=dev)
            o = timeit_ours(lambda: _ext.group_points_grad(g, idx, n), args.iters, flush)
            t = timeit(lambda: ref.group_points_grad(g, idx, n), max(3, args.iters // 3), flush) if ref else None
            rec("group_points_grad", [B, C, n, npoint, ns], o, t, 4 * B * (C * n + npoint * ns + C * npoint * ns))
    if "three_nn" in ops or "interp" in ops:
        for (n, m) in ((512, 256), (1024, 512), (4096, 2048)):      # FP1, FP2, config 5's x4 shape
            if n == 4096:
                big = torch.from_numpy(np.stack([make_scene_xyz(900 + i, 40000) for i in range(B)], 0)).to(dev)
                _, unknown = _ext.furthest_point_sampling_with_xyz(big, 4096)
                known = unknown[:, :2048].contiguous()
            else:
                unknown = levels[3][0] if n == 512 else levels[2][0]
                known = levels[3][1] if n == 512 else levels[2][1]
            o = timeit_ours(lambda: _ext.three_nn(unknown, known), args.iters, flush)
            t = timeit(lambda: ref.three_nn(unknown, known), max(3, args.iters // 3), flush) if ref else None
            rec("three_nn", [B, n, m], o, t, B * (12 * (n + m) + 24 * n))
            d2, idx = _ext.three_nn(unknown, known)
            w = torch.rand(B, n, 3, device=dev)
            feats = torch.randn(B, 256, m, device=dev)
            o = timeit_ours(lambda: _ext.three_interpolate(feats, idx, w), args.iters, flush)
            t = timeit(lambda: ref.three_interpolate(feats, idx, w), max(3, args.iters // 3), flush) if ref else None
            rec("three_interpolate", [B, 256, m, n], o, t, 4 * B * (256 * m + 6 * n + 256 * n))
            g = torch.randn(B, 256, n, device=dev)
            o = timeit_ours(lambda: _ext.three_interpolate_grad(g, idx, w, m), args.iters, flush)
            t = timeit(lambda: ref.three_interpolate_grad(g, idx, w, m), max(3, args.iters // 3), flush) if ref else None
            rec("three_interpolate_grad", [B, 256, n, m], o, t, 4 * B * (256 * n + 6 * n + 256 * m))
    if args.json:
        os.makedirs(os.path.dirname(os.path.abspath(args.json)), exist_ok=True)
        json.dump({"gpu": torch.cuda.get_device_name(0), "rows": out}, open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
