#!/usr/bin/env python
"""Stand-alone duration of the five fused set-abstraction launches of the detector (BASELINE shapes, 8 scenes x 40 k
points), graph-timed (launch overhead excluded, L2 flushed between launches), for min_tiles_per_cta = 0 and 16.

    python tools/time_sa_layers.py [--cf 1] [--json out.json]
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

# name: (npoint, radius, nsample, Cin, (C1, C2, C3))
LAYERS = [("sa1", 2048, 0.2, 64, None, (64, 64, 128)), ("sa2", 1024, 0.4, 32, 128, (128, 128, 256)),
          ("sa3", 512, 0.8, 16, 256, (128, 128, 256)), ("sa4", 256, 1.2, 16, 256, (128, 128, 256))]


def graph_time(fn, flush, reps=8, iters=7):
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
        for _ in range(iters):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            g.replay()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return statistics.median(ts)

    run(ga), run(gb)
    return max(run(ga) - run(gb), 0.0) / reps * 1e3      # microseconds per call


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cf", type=int, default=1, help="raw feature channels of SA1 (1 = config 2, 7 = config 3)")
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    dev = "cuda:0"
    B = args.batch
    torch.manual_seed(0)
    xyz = torch.from_numpy(np.stack([make_scene_xyz(100 + i, 40000) for i in range(B)], 0)).to(dev)
    junk = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    flush = lambda: junk.fill_(1)       # noqa: E731
    rows = []
    feats = torch.randn(B, args.cf, 40000, device=dev) if args.cf else None
    cur_xyz, cur_feat = xyz, feats
    layers = list(LAYERS) + [("vote_agg", 256, 0.3, 16, 256, (128, 128, 128))]
    for name, npoint, radius, ns, cin, (C1, C2, C3) in layers:
        if name == "vote_agg":
            cur_xyz, cur_feat = seeds_xyz, torch.randn(B, 256, seeds_xyz.shape[1], device=dev)   # noqa: F821
        n = cur_xyz.shape[1]
        _, new_xyz = _ext.furthest_point_sampling_with_xyz(cur_xyz, npoint)
        idx = _ext.ball_query(new_xyz, cur_xyz, radius, ns)
        W1 = (torch.randn(C2, C1, device=dev) * 0.1).to(_ext.HALF)
        W2 = (torch.randn(C3, C2, device=dev) * 0.1).to(_ext.HALF)
        b0, b1, b2 = (torch.randn(c, device=dev) * 0.1 for c in (C1, C2, C3))
        if name == "sa1":
            W0 = torch.randn(C1, 3 + args.cf, device=dev) * 0.3
            kw = dict(feat=cur_feat)
        else:
            W0 = torch.randn(C1, 3, device=dev) * 0.3
            kw = dict(G=(torch.randn(B, n, C1, device=dev) * 0.3).to(_ext.HALF))
        rec = {"layer": name, "n": n, "npoint": npoint, "nsample": ns, "widths": [C1, C2, C3],
               "tiles": B * npoint * ns // 128}
        for mt in (0, 16):
            with _ext.launch_options(sa_min_tiles=mt):
                fn = lambda: _ext.sa_fused_forward(cur_xyz, new_xyz, idx, W0, b0, W1, b1, W2, b2, radius=radius,  # noqa: E731
                                                   want_point_major=True, **kw)
                out, _ = fn()
                assert torch.isfinite(out).all()
                rec["us_min_tiles_%d" % mt] = round(graph_time(fn, flush), 2)
        flops = 2.0 * rec["tiles"] * 128 * (C1 * C2 + C2 * C3 + (0 if name != "sa1" else C1 * 16))
        rec["tflops_min_tiles_0"] = round(flops / rec["us_min_tiles_0"] / 1e6, 1)
        rows.append(rec)
        print(json.dumps(rec), flush=True)
        if name == "sa2":
            seeds_xyz = new_xyz       # noqa: F841  (vote aggregation groups the 1024 seeds)
        cur_xyz, cur_feat = new_xyz, None
    print("sum us (min_tiles 0 / 16): %.1f / %.1f" % (sum(r["us_min_tiles_0"] for r in rows),
                                                       sum(r["us_min_tiles_16"] for r in rows)))
    if args.json:
        json.dump({"rows": rows}, open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
