#!/usr/bin/env python
"""One eager detector forward (after 3 warm-up forwards) with the launch hints of the graph pipeline, for ncu:
    ncu --metrics gpu__time_duration.sum,sm__cycles_active.sum,smsp__inst_executed.sum --clock-control none \
        -s <launches of 3 forwards> --csv --log-file gpurun_out/forward.csv python tools/one_forward.py [--config 2]
Prints the number of launches per forward so that -s can be chosen (run once without ncu)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from spacap3d_b200 import _ext, _lib  # noqa: E402

cfg = int(sys.argv[sys.argv.index("--config") + 1]) if "--config" in sys.argv else 2
bench.CFG = dict(bench.CONFIGS[cfg], id=cfg)
bench.n_input_sets = lambda world: 1
model = bench.make_detector(torch.device("cuda", 0))
pc = bench.make_host_batches(0, 1)[0].cuda()
count = [0]
orig = _lib.call


def counting(name, *a):
    count[0] += 1
    return orig(name, *a)


_lib.call = counting
# --default-options: the single-call launch configuration (what bench.py's eager pass and its `roofline` block time)
opts = {} if "--default-options" in sys.argv else dict(fps_algo=_ext.FPS_BUCKET, sa_min_tiles=16, pm_n_tile=256,
                                                         pm_tiles_per_cta=4)     # = GraphedDetector's defaults
with torch.no_grad(), _ext.launch_options(**opts):
    for i in range(4):
        if i == 3:
            torch.cuda.synchronize()
            count[0] = 0
        model({"point_clouds": pc})
torch.cuda.synchronize()
print("C-ABI calls in the last forward:", count[0])
