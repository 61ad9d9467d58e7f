#!/usr/bin/env python
"""Where the end-to-end figure loses against the device-timed one: the same 32-stream pipeline with
(a) device-resident inputs, results stay on the device; (b) inputs from pinned host memory only; (c) results to pinned
host memory only; (d) both (= bench.py's e2e).  Median of 9 regions of --steps submits each."""
import argparse
import json
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--streams", type=int, default=bench.N_STREAMS)
    args = ap.parse_args()
    from spacap3d_b200.pipeline import GraphedDetector
    device = torch.device("cuda:0")
    torch.cuda.set_device(device)
    model = bench.make_detector(device)
    host = bench.make_host_batches(0, 1)
    resident = [h.to(device) for h in host]
    runner = GraphedDetector(model, resident[0], n_streams=args.streams, result_keys=bench.RESULT_KEYS)
    n = len(host)

    def run(src, to_host):
        def submit(i):
            slot = runner._next
            if runner.busy(slot):
                runner.wait(slot)
            runner.submit(src[i % n], to_host=to_host)
        for i in range(40):
            submit(i)
        runner.close()
        cur = torch.cuda.current_stream()
        out = []
        for _ in range(9):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(cur)
            runner.fork_from(e0)
            for k in range(args.steps):
                submit(k)
            runner.join_into(cur)
            e1.record(cur)
            torch.cuda.synchronize()
            out.append(e0.elapsed_time(e1) / args.steps)
            runner.close()
        return statistics.median(out)

    spg = resident[0].shape[0]
    for name, src, th in (("device in, device out", resident, False), ("host in, device out", host, False),
                          ("device in, host out", resident, True), ("host in, host out (e2e)", host, True)):
        ms = run(src, th)
        print(json.dumps({"mode": name, "ms_per_step": round(ms, 4), "scenes_per_s": round(spg / ms * 1e3, 1)}), flush=True)


if __name__ == "__main__":
    main()
