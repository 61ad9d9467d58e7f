"""Seeded inputs for the detection post-processing tests (shared by oracle/make_golden_post.py)."""
import numpy as np


def box_corners(center, size, heading):
    """(…,3),(…,3),(…) -> (…,8,3) float64 in the corner order of utils/box_util.py:360-383 (get_3d_box_batch)."""
    l, w, h = size[..., 0], size[..., 1], size[..., 2]
    x = np.stack([l / 2, l / 2, -l / 2, -l / 2, l / 2, l / 2, -l / 2, -l / 2], -1)
    y = np.stack([h / 2, h / 2, h / 2, h / 2, -h / 2, -h / 2, -h / 2, -h / 2], -1)
    z = np.stack([w / 2, -w / 2, -w / 2, w / 2, w / 2, -w / 2, -w / 2, w / 2], -1)
    c, s = np.cos(heading)[..., None], np.sin(heading)[..., None]
    xr = c * x + s * z
    zr = -s * x + c * z
    return np.stack([xr, y, zr], -1) + center[..., None, :]


def nms_cases():
    """name -> (corners (B,K,8,3) f64, score (B,K) f32 all distinct, cls (B,K) int64, valid (B,K) int32)"""
    out = {}
    for name, (B, K, spread, seed) in {"dense_256": (2, 256, 3.0, 1), "sparse_64": (3, 64, 8.0, 2),
                                       "tiny_5": (1, 5, 1.0, 3)}.items():
        rng = np.random.default_rng(seed)
        center = rng.uniform(-spread, spread, (B, K, 3))
        size = rng.uniform(0.2, 2.0, (B, K, 3))
        heading = np.zeros((B, K)) if name != "sparse_64" else rng.uniform(0, np.pi, (B, K))
        corners = box_corners(center, size, heading).astype(np.float64)
        score = rng.permutation(B * K).reshape(B, K).astype(np.float32) / np.float32(B * K + 1)
        cls = rng.integers(0, 4, (B, K)).astype(np.int64)
        valid = (rng.random((B, K)) < 0.85).astype(np.int32)
        valid[:, 0] = 1
        out[name] = (corners, score, cls, valid)
    return out


def box_cases():
    """name -> (points (B,N,4) f32, corners (B,K,8,3) f64); point coordinates are kept away from box faces."""
    out = {}
    for name, (B, N, K, rot, seed) in {"axis_aligned": (2, 3000, 24, False, 5), "rotated": (1, 2000, 16, True, 6)}.items():
        rng = np.random.default_rng(seed)
        pts = rng.uniform(-3, 3, (B, N, 4)).astype(np.float32)
        center = rng.uniform(-2, 2, (B, K, 3))
        size = rng.uniform(0.05, 2.5, (B, K, 3))
        heading = rng.uniform(0, np.pi, (B, K)) if rot else np.zeros((B, K))
        out[name] = (pts, box_corners(center, size, heading).astype(np.float64))
    return out
