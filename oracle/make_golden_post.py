#!/usr/bin/env python
"""Golden vectors for the detection post-processing oracle (run in the build container, where /root/reference
exists):  python oracle/make_golden_post.py  ->  tests/golden/post_ref.npz

Outputs come from the REFERENCE's own code: utils/nms.py is imported by file path (pure numpy) and the point-in-
box test is scipy's Delaunay(hull).find_simplex(p) >= 0, i.e. data/scannet/model_util_scannet.py:13-22 (that module
itself cannot be imported here: it pulls lib/config.py -> easydict, which is not installed)."""
import importlib.util
import os
import sys

import numpy as np
from scipy.spatial import Delaunay

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases_post  # noqa: E402

import types  # noqa: E402
# utils/nms.py starts with `from utils.pc_utils import bbox_corner_dist_measure` (used only by nms_crnr_dist);
# pc_utils needs plyfile / trimesh / matplotlib, which are not installed here, so that one name is stubbed.
_pkg, _pc = types.ModuleType("utils"), types.ModuleType("utils.pc_utils")
_pc.bbox_corner_dist_measure = None
sys.modules.setdefault("utils", _pkg)
sys.modules["utils.pc_utils"] = _pc
spec = importlib.util.spec_from_file_location("ref_nms", "/root/reference/utils/nms.py")
ref_nms = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref_nms)


def ref_pick(corners, score, cls, valid, mode, old_type, thr):
    """The per-scene loops of ap_helper.py:82-137 around the reference's nms functions."""
    B, K = corners.shape[:2]
    mask = np.zeros((B, K), np.int32)
    for i in range(B):
        lo, hi = corners[i].min(1), corners[i].max(1)
        if mode == 0:
            boxes = np.stack([lo[:, 0], lo[:, 2], hi[:, 0], hi[:, 2], score[i]], 1).astype(np.float64)
            fn = ref_nms.nms_2d_faster
        elif mode == 1:
            boxes = np.concatenate([lo, hi, score[i][:, None]], 1).astype(np.float64)
            fn = ref_nms.nms_3d_faster
        else:
            boxes = np.concatenate([lo, hi, score[i][:, None], cls[i][:, None]], 1).astype(np.float64)
            fn = ref_nms.nms_3d_faster_samecls
        ids = np.where(valid[i] == 1)[0]
        with np.errstate(divide="ignore", invalid="ignore"):
            pick = fn(boxes[valid[i] == 1, :], thr, old_type)
        mask[i, ids[pick]] = 1
    return mask


def main():
    out = {}
    for name, (corners, score, cls, valid) in cases_post.nms_cases().items():
        for mode in (0, 1, 2):
            for old in (False, True):
                out["nms/%s/m%d_o%d" % (name, mode, int(old))] = ref_pick(corners, score, cls, valid, mode, old, 0.25)
    for name, (pts, corners) in cases_post.box_cases().items():
        B, K = corners.shape[:2]
        cnt = np.zeros((B, K), np.int32)
        for b in range(B):
            for k in range(K):
                cnt[b, k] = int((Delaunay(corners[b, k]).find_simplex(pts[b, :, :3]) >= 0).sum())
        out["box/%s/count" % name] = cnt
    path = os.path.join(ROOT, "tests", "golden", "post_ref.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, len(out), "arrays")


if __name__ == "__main__":
    main()
