#!/usr/bin/env python
"""Follow-up to tools/e2e_probe.py: what about the host -> device copy costs the pipeline 49 us per step?"""
import json
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    from spacap3d_b200.pipeline import GraphedDetector
    device = torch.device("cuda:0")
    torch.cuda.set_device(device)
    model = bench.make_detector(device)
    host = bench.make_host_batches(0, 1)
    resident = [h.to(device) for h in host]
    print(json.dumps({"pinned": host[0].is_pinned(), "bytes": host[0].numel() * 4}))
    # (1) the copy alone
    s = torch.cuda.Stream()
    dst = torch.empty_like(resident[0])
    with torch.cuda.stream(s):
        for _ in range(10):
            dst.copy_(host[0], non_blocking=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for i in range(100):
            dst.copy_(host[i % len(host)], non_blocking=True)
        e1.record(s)
    torch.cuda.synchronize()
    print(json.dumps({"h2d_alone_us_per_copy": round(e0.elapsed_time(e1) * 10, 1)}))
    runner = GraphedDetector(model, resident[0], n_streams=bench.N_STREAMS, result_keys=bench.RESULT_KEYS)
    n = len(host)
    copy_stream = torch.cuda.Stream()
    staged = [torch.empty_like(resident[0]) for _ in range(runner.n)]
    evs = [torch.cuda.Event() for _ in range(runner.n)]
    frees = [torch.cuda.Event() for _ in range(runner.n)]

    def submit_plain(i, frac=1.0):
        slot = runner._next
        runner._next = (slot + 1) % runner.n
        st = runner.streams[slot]
        with torch.cuda.stream(st):
            if frac >= 1.0:
                runner.static_in[slot].copy_(host[i % n], non_blocking=True)
            else:
                k = int(host[0].shape[0] * frac)
                runner.static_in[slot][:k].copy_(host[i % n][:k], non_blocking=True)
                runner.static_in[slot][k:].copy_(resident[i % n][k:], non_blocking=True)
            runner.graphs[slot].replay()

    def submit_copy_stream(i):
        slot = runner._next
        runner._next = (slot + 1) % runner.n
        st = runner.streams[slot]
        copy_stream.wait_event(frees[slot])                 # the slot's previous graph has read its staging buffer
        with torch.cuda.stream(copy_stream):
            staged[slot].copy_(host[i % n], non_blocking=True)
            evs[slot].record(copy_stream)
        st.wait_event(evs[slot])
        with torch.cuda.stream(st):
            runner.static_in[slot].copy_(staged[slot], non_blocking=True)
            runner.graphs[slot].replay()
            frees[slot].record(st)

    def measure(submit, steps=100):
        for i in range(40):
            submit(i)
        runner.wait_all()
        cur = torch.cuda.current_stream()
        out = []
        for _ in range(7):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(cur)
            runner.fork_from(e0)
            for k in range(steps):
                submit(k)
            runner.join_into(cur)
            copy_stream.synchronize()
            e1.record(cur)
            torch.cuda.synchronize()
            out.append(e0.elapsed_time(e1) / steps)
        return round(statistics.median(out), 4)

    print(json.dumps({"host in, same stream": measure(submit_plain)}), flush=True)
    print(json.dumps({"half of the batch from the host": measure(lambda i: submit_plain(i, 0.5))}), flush=True)
    print(json.dumps({"host in via one copy stream + staging": measure(submit_copy_stream)}), flush=True)
    print(json.dumps({"device in": measure(lambda i: runner.submit(resident[i % n]))}), flush=True)


if __name__ == "__main__":
    main()
