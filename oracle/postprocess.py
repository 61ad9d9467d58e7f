"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): numpy restatement of the reference's detection
post-processing, the checker for spacap3d_b200/postprocess.py and csrc/postprocess.cu.

Follows /root/reference
  lib/ap_helper.py:44-160                  parse_predictions
  utils/nms.py:39-70, 72-107, 109-147      nms_2d_faster, nms_3d_faster, nms_3d_faster_samecls
  data/scannet/model_util_scannet.py:13-22 in_hull / extract_pc_in_box3d (scipy Delaunay hull membership)
Pinned by tests/golden/post_ref.npz, produced by oracle/make_golden_post.py from the reference's own
utils/nms.py (imported by path) and scipy's Delaunay.find_simplex, the call the reference makes.
Differences by construction: equal scores are ordered by a stable sort (numpy's default introsort is not
specified there); box membership is evaluated with three edge projections instead of a Delaunay triangulation
(identical except for points within ~1e-12 of a face).
"""
import numpy as np


def softmax(x):
    """ap_helper.py:37-42"""
    probs = np.exp(x - np.max(x, axis=-1, keepdims=True))
    probs /= np.sum(probs, axis=-1, keepdims=True)
    return probs


def box_point_counts(points, corners):
    """points (B,N,>=3) f32, corners (B,K,8,3) f64 in get_3d_box_batch order -> (B,K) int32 number of points
    inside each box (edges from corner 0 to corners 1, 3, 4)."""
    B, K = corners.shape[:2]
    out = np.zeros((B, K), np.int32)
    for b in range(B):
        p = points[b, :, :3].astype(np.float64)
        for k in range(K):
            c = corners[b, k]
            inside = np.ones(len(p), bool)
            d = p - c[0]
            for o in (1, 3, 4):
                e = c[o] - c[0]
                n2 = float(e @ e)
                t = (d[:, 0] * e[0] + d[:, 1] * e[1] + d[:, 2] * e[2]) * (1.0 / n2 if n2 > 0 else 0.0)
                inside &= (t >= 0.0) & (t <= 1.0)
            out[b, k] = int(inside.sum())
    return out


def nms_boxes(corners, score, cls, valid, mode, old_type, thr):
    """corners (B,K,8,3) f64, score (B,K) f32, cls (B,K) int or None, valid (B,K) bool/int or None ->
    pick mask (B,K) int32.  mode 0/1/2 = nms_2d_faster / nms_3d_faster / nms_3d_faster_samecls."""
    B, K = corners.shape[:2]
    pick = np.zeros((B, K), np.int32)
    for b in range(B):
        lo, hi = corners[b].min(1), corners[b].max(1)                      # (K,3)
        sc = score[b].astype(np.float64)
        keep = np.ones(K, bool) if valid is None else (np.asarray(valid[b]) != 0)
        ids = np.nonzero(keep)[0]
        if mode == 0:
            x1, y1, x2, y2 = lo[ids, 0], lo[ids, 2], hi[ids, 0], hi[ids, 2]
            area = (x2 - x1) * (y2 - y1)
        else:
            x1, y1, z1, x2, y2, z2 = lo[ids, 0], lo[ids, 1], lo[ids, 2], hi[ids, 0], hi[ids, 1], hi[ids, 2]
            area = (x2 - x1) * (y2 - y1) * (z2 - z1)
        c = None if cls is None else np.asarray(cls[b])[ids]
        I = np.argsort(sc[ids], kind="stable")
        with np.errstate(divide="ignore", invalid="ignore"):
            while I.size != 0:
                last = I.size
                i = I[-1]
                pick[b, ids[i]] = 1
                rest = I[:last - 1]
                xx1, xx2 = np.maximum(x1[i], x1[rest]), np.minimum(x2[i], x2[rest])
                yy1, yy2 = np.maximum(y1[i], y1[rest]), np.minimum(y2[i], y2[rest])
                if mode == 0:
                    inter = np.maximum(0, xx2 - xx1) * np.maximum(0, yy2 - yy1)
                    o = inter / area[rest] if old_type else inter / (area[i] + area[rest] - inter)
                else:
                    zz1, zz2 = np.maximum(z1[i], z1[rest]), np.minimum(z2[i], z2[rest])
                    inter = np.maximum(0, xx2 - xx1) * np.maximum(0, yy2 - yy1) * np.maximum(0, zz2 - zz1)
                    if old_type:
                        o = inter / area[rest]
                    elif mode == 2:
                        o = inter / (area[i] + area[rest] - inter + 1e-8)
                    else:
                        o = inter / (area[i] + area[rest] - inter)
                    if mode == 2:
                        o = o * (c[i] == c[rest])
                I = np.delete(I, np.concatenate(([last - 1], np.where(o > thr)[0])))
    return pick


def parse_predictions(end_points, config_dict):
    """end_points: numpy arrays 'point_clouds' (B,N,>=3), 'bbox_corner' (B,K,8,3) f64, 'objectness_scores' (B,K,2),
    'sem_cls_scores' (B,K,C), 'sem_cls' (B,K).  Returns (batch_pred_map_cls, pred_mask) like ap_helper.py:44-160."""
    corners = end_points["bbox_corner"]
    B, K = corners.shape[:2]
    sem_cls_probs = softmax(end_points["sem_cls_scores"])
    obj_prob = softmax(end_points["objectness_scores"])[:, :, 1]
    nonempty = np.ones((B, K), bool)
    if config_dict["remove_empty_box"]:
        nonempty = box_point_counts(end_points["point_clouds"], corners) >= 5
    if not config_dict["use_3d_nms"]:
        mode = 0
    elif not config_dict["cls_nms"]:
        mode = 1
    else:
        mode = 2
    pred_mask = nms_boxes(corners, obj_prob, end_points["sem_cls"], nonempty, mode,
                          bool(config_dict["use_old_type_nms"]), config_dict["nms_iou"])
    out = []
    for i in range(B):
        if config_dict["per_class_proposal"]:
            cur = []
            for ii in range(sem_cls_probs.shape[2]):
                cur += [(ii, corners[i, j], sem_cls_probs[i, j, ii] * obj_prob[i, j]) for j in range(K)
                        if pred_mask[i, j] == 1 and obj_prob[i, j] > config_dict["conf_thresh"]]
            out.append(cur)
        else:
            out.append([(int(end_points["sem_cls"][i, j]), corners[i, j], obj_prob[i, j]) for j in range(K)
                        if pred_mask[i, j] == 1 and obj_prob[i, j] > config_dict["conf_thresh"]])
    return out, pred_mask
