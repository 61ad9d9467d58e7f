#!/usr/bin/env python
"""BASELINE config 5: detector forward + backward (+ DDP-style gradient all-reduce) at growing cloud sizes.

    python tools/bench_train.py [--points 40000,80000,120000,200000] [--scenes 4] [--iters 5] [--ref]
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/bench_train.py

Training mode: autograd through the unfused sm_100a kernels (FPS, ball query, grouping, three_nn,
interpolate and their backward passes) + cuDNN convolutions, exactly the reference's op sequence
(pointnet2_modules.py:227-276).  The loss is a surrogate (mean square of every head output): the
reference's loss (lib/loss_helper.py) is outside the hot path.  SA npoint is scaled with the cloud
(x1, x2, x3, x5 -> 2048..8192+).  `--ref` also times the reference's own CUDA
extension (oracle/_ref) driving the same modules on rank 0.  One JSON line per cloud size.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def build_model(scale, device):
    from spacap3d_b200.detector import VoteNetDetector
    torch.manual_seed(0)
    model = VoteNetDetector(input_feature_dim=1)
    bb = model.backbone_net
    for sa, base in ((bb.sa1, 2048), (bb.sa2, 1024), (bb.sa3, 512), (bb.sa4, 256)):
        sa.npoint = base * scale
    return model.to(device).train()


def surrogate_loss(out):
    """Mean square of every head output (the reference's loss, lib/loss_helper.py, is outside the hot path)."""
    return sum((out[k].float() ** 2).mean() for k in ("objectness_scores", "center", "sem_cls_scores",
                                                      "size_scores", "size_residuals", "vote_xyz"))


def step(model, pc, world):
    out = model({"point_clouds": pc})
    loss = surrogate_loss(out)
    loss.backward()
    if world > 1:
        from spacap3d_b200.dist import allreduce_gradients
        allreduce_gradients(model)
    return loss


def time_steps(fn, iters, warmup=2):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in ev:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    return float(np.median([a.elapsed_time(b) for a, b in ev]))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", default="40000,80000,120000,200000")
    ap.add_argument("--scenes", type=int, default=4, help="scenes per GPU")
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--ref", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)
    from spacap3d_b200.dist import max_over_ranks
    from spacap3d_b200.scenes import make_scene
    ref = None
    if args.ref and rank == 0:
        from oracle.build_ref import load_ref
        ref = load_ref()
    for n in [int(x) for x in args.points.split(",")]:
        scale = max(1, round(n / 40000))
        model = build_model(scale, dev)
        pc = torch.from_numpy(np.stack([make_scene(5000 + rank * 100 + i, n) for i in range(args.scenes)], 0)).to(dev)

        def ours():
            model.zero_grad(set_to_none=True)
            step(model, pc, world)

        ms = time_steps(ours, args.iters)
        ms = max_over_ranks(ms, device=dev) if world > 1 else ms
        rec = {"config": "fwd+bwd, %d scenes/GPU x %d pts, SA npoint x%d" % (args.scenes, n, scale), "n_gpus": world,
               "ms_per_step": round(ms, 3), "scenes_per_s": round(world * args.scenes / ms * 1e3, 2)}
        if world == 1 and not args.no_graph:
            # the step has no host synchronisation (box decode stays on the device), so forward + backward can be
            # captured once and replayed: removes the ~3 ms/step of Python/launch overhead of ~600 small launches
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                ours()
            torch.cuda.current_stream().wait_stream(side)
            model.zero_grad(set_to_none=True)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                step(model, pc, 1)
            gms = time_steps(graph.replay, args.iters)
            rec["graphed_ms_per_step"] = round(gms, 3)
            rec["graphed_scenes_per_s"] = round(args.scenes / gms * 1e3, 2)
            del graph
        if ref is not None:
            import bench

            from spacap3d_b200 import pytorch_utils

            def theirs():      # reference extension + the plain cuDNN BatchNorm / ReLU / max_pool2d sequence
                model.zero_grad(set_to_none=True)
                pytorch_utils.FUSED_BN_RELU_TRAINING = False
                try:
                    with bench.swapped_ops(ref, host_decode=False):
                        step(model, pc, 1)
                finally:
                    pytorch_utils.FUSED_BN_RELU_TRAINING = True

            try:
                rms = time_steps(theirs, max(2, args.iters // 2), warmup=1)
                rec["reference_ext_ms_per_step"] = round(rms, 3)
                rec["speedup_vs_reference_ext"] = round(rms / ms, 2)
            except Exception as e:   # the reference FPS kernel has a fixed 512-thread block; very large N still works, but be safe
                rec["reference_ext_error"] = str(e)[:200]
        if rank == 0:
            print(json.dumps(rec), flush=True)
        del model, pc
        torch.cuda.empty_cache()
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
