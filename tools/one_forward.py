import sys, numpy as np, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
model = bench.make_detector(torch.device("cuda", 0))
bench.N_INPUT_SETS = 1
host = bench.make_host_batches(0)
pc = host[0].cuda()
with torch.no_grad():
    for _ in range(3):
        model({"point_clouds": pc})
torch.cuda.synchronize()
