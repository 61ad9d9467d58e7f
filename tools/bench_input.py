#!/usr/bin/env python
"""time the assembly of one training batch (8 items x 40 k of 50 k vertices, augmentation + vote
labels + box augmentation) -- device kernels (scenes resident in HBM) vs the numpy restatement of the reference's
__getitem__ point code (oracle/input_pipeline.py) on the host.  One JSON line per feature configuration."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import cases_input
    from oracle import input_pipeline as oi
    from spacap3d_b200 import _ext, input_pipeline as ip
    peak = 6555.8
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    B, M, P = 8, 50000, 40000
    st = ip.DeviceSceneStore("cuda")
    scenes = []
    for s in range(B):
        v, inst, sem, bb, mv = cases_input.make_scene(M, 60, 200 + s, multiview=True)
        scenes.append((v, inst, sem, bb, mv))
        st.add_scene("s%d" % s, v, inst, sem, bb, mv)
    st.finalize()
    ids = ["s%d" % s for s in range(B)]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for name, kw in (("xyz+height", dict(use_height=True)),
                     ("xyz+rgb+normal+height", dict(use_color=True, use_normal=True, use_height=True)),
                     ("xyz+multiview+normal+height", dict(use_normal=True, use_multiview=True, use_height=True))):
        t0 = time.perf_counter()
        draws = [ip.draw_item(np.random.RandomState(s), M, P, True) for s in range(B)]
        draw_ms = (time.perf_counter() - t0) * 1e3
        out = st.make_batch(ids, draws, **kw)
        C = out["point_clouds"].shape[2]
        n_mv = 128 if kw.get("use_multiview") else 0
        # component timing on prebuilt device arguments (L2 flushed before every iteration)
        choices = out["choices"]
        aug = torch.from_numpy(np.stack([d[1] for d in draws])).cuda()
        row0 = torch.from_numpy(st.row0).cuda()
        fh = st.floor_height
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        acc = np.zeros(3)
        iters = 20
        for it in range(iters + 3):
            flush.zero_()
            e[0].record()
            pc = _ext.prepare_point_clouds(st.verts, row0, choices, multiview=st.multiview if n_mv else None,
                                           floor_height=fh, aug=aug, mean_rgb=ip.MEAN_COLOR_RGB,
                                           use_color=kw.get("use_color", False), use_normal=kw.get("use_normal", False))
            e[1].record()
            _ext.vote_labels(pc, st.instance_labels, st.semantic_labels, row0, choices, st.max_instances,
                             ip.sem_mask_of())
            e[2].record()
            _ext.augment_boxes(st.boxes, aug)
            e[3].record()
            torch.cuda.synchronize()
            if it >= 3:
                acc += [e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2]), e[2].elapsed_time(e[3])]
        prep_ms, vote_ms, box_ms = acc / iters
        # whole call incl. the host->device copies of the draws
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(10):
            st.make_batch(ids, draws, **kw)
        torch.cuda.synchronize()
        call_ms = (time.perf_counter() - t0) * 1e3 / 10
        # host: numpy port of the reference's per-item work (one DataLoader worker), excluding the RNG draws
        t0 = time.perf_counter()
        n_cpu = 2
        for s in range(n_cpu):
            v, inst, sem, bb, mv = scenes[s]
            fhs = oi.floor_height(v[:, 2])
            want = oi.prepare_point_cloud(v, draws[s][0], mv if n_mv else None, fhs, draws[s][1],
                                          kw.get("use_color", False), kw.get("use_normal", False))
            oi.vote_labels(want, inst[draws[s][0]], sem[draws[s][0]])
            oi.augment_boxes(np.zeros((128, 6)), draws[s][1])
        cpu_ms = (time.perf_counter() - t0) * 1e3 / n_cpu * B
        assert np.array_equal(out["point_clouds"][n_cpu - 1].cpu().numpy(), want)
        src_row = 4 * (3 + (3 if kw.get("use_color") else 0) + (3 if kw.get("use_normal") else 0)) + 4 * n_mv
        algo = B * P * (4 * C + src_row + 4)
        print(json.dumps({"op": "input_pipeline", "features": name, "B": B, "vertices": M, "points": P, "channels": C,
                          "prepare_ms": round(float(prep_ms), 4), "vote_labels_ms": round(float(vote_ms), 4),
                          "augment_boxes_ms": round(float(box_ms), 4),
                          "make_batch_call_ms": round(call_ms, 3), "host_rng_draws_ms": round(draw_ms, 2),
                          "numpy_port_ms_per_batch_1_worker": round(cpu_ms, 1),
                          "prepare_algo_bytes": algo, "prepare_gbs": round(algo / prep_ms / 1e6, 1),
                          "prepare_frac_of_hbm_peak": round(algo / prep_ms / 1e6 / peak, 3), "hbm_peak_gbs": peak,
                          "speedup_vs_numpy_port": round(cpu_ms / float(prep_ms + vote_ms + box_ms), 1)}))


if __name__ == "__main__":
    main()
