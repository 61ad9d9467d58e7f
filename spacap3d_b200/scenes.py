"""Synthetic "ScanNet-shaped" scenes -- no dataset is available offline.

A scene is N points on the surfaces of an 8 m x 6 m x 3 m room (floor-heavy) plus 10-30
axis-aligned boxes standing on the floor, Gaussian jitter sigma = 1 cm, translated so the room
centre sits at the origin (this makes the reference's "|p|^2 <= 1e-3 is skipped" quirk fire,
sampling_gpu.cu:100-101).  A fraction of scenes is built from fewer distinct points resampled
with replacement (the reference does this for small scans, utils/pc_utils.py:32-36), which
creates exact duplicate points and therefore exact distance ties in FPS.

Feature channels follow lib/dataset.py:309-333: rgb, normal, multiview, height (in that order).
Everything is numpy + a seeded Generator so the same seed gives the same bytes on every host.
"""
import numpy as np

ROOM = np.array([8.0, 6.0, 3.0], dtype=np.float64)


def _sample_box_surface(rng, n, lo, hi, face_weights=None):
    """n points uniformly on the faces of the axis-aligned box [lo, hi]."""
    lo = np.asarray(lo, np.float64)
    hi = np.asarray(hi, np.float64)
    ext = hi - lo
    areas = np.array([ext[1] * ext[2], ext[1] * ext[2], ext[0] * ext[2], ext[0] * ext[2],
                      ext[0] * ext[1], ext[0] * ext[1]])
    if face_weights is not None:
        areas = areas * np.asarray(face_weights, np.float64)
    face = rng.choice(6, size=n, p=areas / areas.sum())
    pts = lo + rng.random((n, 3)) * ext
    axis = face // 2
    side = face % 2
    pts[np.arange(n), axis] = np.where(side == 0, lo[axis], hi[axis])
    normals = np.zeros((n, 3), np.float64)
    normals[np.arange(n), axis] = np.where(side == 0, -1.0, 1.0)
    return pts, normals


def make_scene_xyz(seed, n_points=40000, with_replacement=None, return_normals=False):
    """(n_points, 3) float32 room-shaped cloud centred at the origin."""
    rng = np.random.default_rng(seed)
    if with_replacement is None:
        with_replacement = rng.random() < 0.05
    n_distinct = int(n_points * 0.75) if with_replacement else n_points
    n_boxes = int(rng.integers(10, 31))
    n_room = int(n_distinct * 0.55)
    per_box = (n_distinct - n_room) // n_boxes
    # room shell: floor x3 weight, ceiling x0.5
    pts, nrm = _sample_box_surface(rng, n_room + (n_distinct - n_room - per_box * n_boxes),
                                   -ROOM / 2, ROOM / 2, [1, 1, 1, 1, 3.0, 0.5])
    chunks, nchunks = [pts], [-nrm]
    for _ in range(n_boxes):
        size = rng.uniform(0.3, 2.0, 3)
        size[2] = min(size[2], 2.2)
        cx = rng.uniform(-ROOM[0] / 2 + size[0] / 2, ROOM[0] / 2 - size[0] / 2)
        cy = rng.uniform(-ROOM[1] / 2 + size[1] / 2, ROOM[1] / 2 - size[1] / 2)
        lo = np.array([cx - size[0] / 2, cy - size[1] / 2, -ROOM[2] / 2])
        hi = lo + size
        p, nm = _sample_box_surface(rng, per_box, lo, hi, [1, 1, 1, 1, 0.0, 1.0])
        chunks.append(p)
        nchunks.append(nm)
    xyz = np.concatenate(chunks, 0)
    normals = np.concatenate(nchunks, 0)
    xyz = xyz + rng.normal(0.0, 0.01, xyz.shape)
    perm = rng.permutation(xyz.shape[0])
    xyz, normals = xyz[perm], normals[perm]
    if with_replacement:
        pick = rng.integers(0, xyz.shape[0], n_points)
        xyz, normals = xyz[pick], normals[pick]
    xyz = xyz.astype(np.float32)
    if return_normals:
        return xyz, normals.astype(np.float32)
    return xyz


def make_scene(seed, n_points=40000, use_color=False, use_normal=False, use_multiview=False,
               use_height=True, with_replacement=None):
    """(n_points, 3 + C) float32 cloud in the reference's channel order
    xyz | rgb(3) | normal(3) | multiview(128) | height(1)   (lib/dataset.py:309-333)."""
    xyz, normals = make_scene_xyz(seed, n_points, with_replacement, return_normals=True)
    rng = np.random.default_rng(seed + 7919)
    cols = [xyz]
    if use_color:
        rgb = (rng.uniform(0, 255, (n_points, 3)) - np.array([109.8, 97.2, 83.8])) / 256.0
        cols.append(rgb.astype(np.float32))
    if use_normal:
        cols.append(normals)
    if use_multiview:
        mv = np.maximum(rng.standard_normal((n_points, 128)), 0.0).astype(np.float32)
        mv[rng.random(n_points) < 0.2] = 0.0
        cols.append(mv)
    if use_height:
        floor = np.percentile(xyz[:, 2], 0.99)
        cols.append((xyz[:, 2] - floor)[:, None].astype(np.float32))
    return np.concatenate(cols, 1).astype(np.float32)


def make_batch(config_id, batch, n_points=40000, **kw):
    """(batch, n_points, 3+C) float32; scene seed = 1000*config_id + scene_index."""
    return np.stack([make_scene(1000 * config_id + i, n_points, **kw) for i in range(batch)], 0)
