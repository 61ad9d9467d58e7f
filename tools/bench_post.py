#!/usr/bin/env python
"""time parse_predictions (8 scenes x 40k points x 256 proposals, SpaCap3D's eval settings) --
device kernels vs the numpy restatement of the reference's host code (oracle/postprocess.py; the reference itself
uses a scipy Delaunay hull test per box, which is slower still).  Prints one JSON line."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import bench
    from oracle import postprocess as op
    from spacap3d_b200.postprocess import parse_predictions, predictions_mask
    dev = torch.device("cuda", 0)
    model = bench.make_detector(dev)
    bench.N_INPUT_SETS = 1
    pc = bench.make_host_batches(0)[0].to(dev)
    with torch.no_grad():
        out = model({"point_clouds": pc})
    cfg = {"remove_empty_box": True, "use_3d_nms": True, "nms_iou": 0.25, "use_old_type_nms": False, "cls_nms": True,
           "per_class_proposal": True, "conf_thresh": 0.05}
    for _ in range(3):
        predictions_mask(out, cfg)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        predictions_mask(out, cfg)
    e1.record()
    torch.cuda.synchronize()
    dev_ms = e0.elapsed_time(e1) / 20
    t0 = time.perf_counter()
    parse_predictions(out, cfg)
    full_ms = (time.perf_counter() - t0) * 1e3
    ref_in = {k: out[k].detach().cpu().numpy() for k in ("point_clouds", "bbox_corner", "objectness_scores",
                                                         "sem_cls_scores", "sem_cls")}
    t0 = time.perf_counter()
    op.parse_predictions(ref_in, cfg)
    cpu_ms = (time.perf_counter() - t0) * 1e3
    print(json.dumps({"op": "parse_predictions", "shape": [int(pc.shape[0]), int(pc.shape[1]), 256],
                      "device_mask_ms": round(dev_ms, 4), "device_incl_host_lists_ms": round(full_ms, 3),
                      "numpy_port_ms": round(cpu_ms, 1), "speedup_mask": round(cpu_ms / dev_ms, 1)}))


if __name__ == "__main__":
    main()
