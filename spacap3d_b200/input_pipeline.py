"""Device-resident input pipeline: what `ScannetReferenceDataset.__getitem__` + the DataLoader's
collate produce for the detector (lib/dataset.py:291-531), assembled on the GPU from scenes kept in HBM.

    store = DeviceSceneStore(device)                       # once
    store.add_scene("scene0000_00", mesh_vertices, instance_labels, semantic_labels, instance_bboxes, multiview)
    store.finalize()                                       # packs the tables, computes the floor heights
    batch = store.make_batch(scene_ids, draws, use_color=..., use_normal=..., use_multiview=..., use_height=...)

`draws` are the per-item random numbers, drawn ON THE HOST in the reference's exact np.random call order
(`draw_item`), so seeding numpy the same way yields the batch the reference's dataset would have produced.
The host does no per-point work; everything O(points) runs in csrc/input_pipeline.cu.

Keys returned (same names / dtypes / shapes as the reference's data_dict after default collate):
  point_clouds (B,P,C) f32, vote_label (B,P,9) f32, vote_label_mask (B,P) i64, center_label (B,128,3) f32,
  box_label_mask (B,128) f32, num_bbox (B,) i64, plus target_bboxes (B,128,6) f64 (the augmented boxes the remaining
  label code of lib/dataset.py:433-470 derives its size residuals / corners from) and choices (B,P) i32.
Language / reference-object labels depend on the caption annotation, not on the point data, and stay with the caller.
"""
import numpy as np
import torch

from . import _ext

MAX_NUM_OBJ = 128                                   # lib/dataset.py:27
MEAN_COLOR_RGB = (109.8, 97.2, 83.8)                # lib/dataset.py:28
NYU40IDS = (3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 23, 24, 25, 26, 27, 28, 29, 30, 31,
            32, 33, 34, 35, 36, 37, 38, 39, 40)     # data/scannet/model_util_scannet.py:88
_TRANSLATIONS = np.arange(-0.5, 0.501, 0.001)       # lib/dataset.py:234


def sem_mask_of(ids=NYU40IDS):
    m = 0
    for i in ids:
        m |= 1 << int(i)
    return m


def _rotx(t):                                       # utils/pc_utils.py:282-288
    c, s = np.cos(t), np.sin(t)
    return np.array([[1, 0, 0], [0, c, -s], [0, s, c]])


def _roty(t):                                       # utils/pc_utils.py:290-296
    c, s = np.cos(t), np.sin(t)
    return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])


def _rotz(t):                                       # utils/pc_utils.py:314-320
    c, s = np.cos(t), np.sin(t)
    return np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]])


def draw_item(rs, num_vertices, num_points, augment):
    """The random draws of one __getitem__, in the reference's order: the subsample (utils/pc_utils.py:32-40 via
    lib/dataset.py:335), two flips, three angles (lib/dataset.py:367-401), three translation factors
    (lib/dataset.py:234-236).  `rs` is `np.random` (the reference's global stream) or a RandomState.
    Returns (choices int64 (num_points,), aug float64 (32,) or None)."""
    choices = rs.choice(num_vertices, num_points, replace=(num_vertices < num_points))
    if not augment:
        return choices, None
    aug = np.zeros(32, np.float64)
    aug[0] = 1.0 if rs.random() > 0.5 else 0.0
    aug[1] = 1.0 if rs.random() > 0.5 else 0.0
    for k, rot in enumerate((_rotx, _roty, _rotz)):
        angle = (rs.random() * np.pi / 18) - np.pi / 36
        aug[2 + 9 * k:11 + 9 * k] = rot(angle).reshape(-1)
    for k in range(3):
        aug[29 + k] = rs.choice(_TRANSLATIONS, size=1)[0]
    return choices, aug


def draw_batch_device(counts, num_points, generator=None, device="cuda"):
    """Throughput alternative to `draw_item` for the subsample only: uniform random subsets (without replacement where
    the scene has enough vertices) drawn ON THE DEVICE with torch's generator -- the same distribution as
    np.random.choice (utils/pc_utils.py:36) but not the same stream; np.random.choice of 40 k out of 50 k costs ~0.7 ms
    per item on the host, which would cap a single feeder thread at ~1.4 k scenes/s.  counts: per-item vertex counts.
    Returns choices (B,num_points) int32 on the device."""
    counts = [int(c) for c in counts]
    M = max(counts)
    cnt = torch.tensor(counts, device=device)
    keys = torch.rand((len(counts), M), generator=generator, device=device)
    keys.masked_fill_(torch.arange(M, device=device)[None, :] >= cnt[:, None], 2.0)    # rows beyond the scene sort last
    perm = keys.argsort(dim=1)
    if M < num_points:                                                                # every scene is too small
        perm = torch.nn.functional.pad(perm, (0, num_points - M))
    choices = perm[:, :num_points]
    if min(counts) < num_points:                                                      # replace=True for small scenes
        with_rep = (torch.rand((len(counts), num_points), generator=generator, device=device) * cnt[:, None]).long()
        with_rep = torch.minimum(with_rep, cnt[:, None] - 1)
        choices = torch.where((cnt < num_points)[:, None], with_rep, choices)
    return choices.to(torch.int32).contiguous()


class DeviceSceneStore:
    """All scenes of a split packed into device tables (see csrc/input_pipeline.cu for the sizing argument)."""

    def __init__(self, device="cuda"):
        self.device = torch.device(device)
        self._pending = []
        self.index = {}
        self.verts = None

    def add_scene(self, scene_id, mesh_vertices, instance_labels, semantic_labels, instance_bboxes, multiview=None):
        mv = None if multiview is None else np.ascontiguousarray(multiview, np.float32)
        self._pending.append((scene_id, np.ascontiguousarray(mesh_vertices, np.float32),
                              np.asarray(instance_labels).astype(np.int32), np.asarray(semantic_labels).astype(np.int32),
                              np.asarray(instance_bboxes, np.float64), mv))

    def finalize(self):
        dev = self.device
        rows, n = [], 0
        for sid, v, il, sl, bb, mv in self._pending:
            self.index[sid] = len(rows)
            rows.append((n, v.shape[0]))
            n += v.shape[0]
        self.row0 = np.array([r[0] for r in rows], np.int64)
        self.count = np.array([r[1] for r in rows], np.int64)
        self.verts = torch.from_numpy(np.concatenate([p[1] for p in self._pending], 0)).to(dev)
        self.instance_labels = torch.from_numpy(np.concatenate([p[2] for p in self._pending], 0)).to(dev)
        self.semantic_labels = torch.from_numpy(np.concatenate([p[3] for p in self._pending], 0)).to(dev)
        self.max_instances = int(self.instance_labels.max().item()) + 1 if n else 1
        has_mv = [p[5] is not None for p in self._pending]
        self.multiview = (torch.from_numpy(np.concatenate([p[5] for p in self._pending], 0)).to(dev)
                          if all(has_mv) and has_mv else None)
        S = len(rows)
        boxes = np.zeros((S, MAX_NUM_OBJ, 6), np.float64)
        self.num_bbox = np.zeros(S, np.int64)
        for i, p in enumerate(self._pending):
            nb = min(p[4].shape[0], MAX_NUM_OBJ)                      # lib/dataset.py:361-363
            boxes[i, :nb] = p[4][:MAX_NUM_OBJ, 0:6]
            self.num_bbox[i] = nb
        self.boxes = torch.from_numpy(boxes).to(dev)
        fh = torch.empty(S, dtype=torch.float32, device=dev)
        for i, (r0, m) in enumerate(rows):                            # once per scene (lib/dataset.py:331)
            fh[i:i + 1] = _ext.scene_floor_height(self.verts[r0:r0 + m])
        self.floor_height = fh
        self._pending = []
        return self

    def make_batch(self, scene_ids, draws, use_color=False, use_normal=False, use_multiview=False, use_height=True,
                   want_votes=True):
        dev = self.device
        if use_multiview and self.multiview is None:
            raise RuntimeError("use_multiview: not every scene of this store was added with multiview features")
        sidx = np.array([self.index[s] for s in scene_ids], np.int64)
        if isinstance(draws, tuple):                       # (choices (B,P) int32 device tensor, aug (B,32) ndarray or None)
            choices, aug_np = draws
            augment = aug_np is not None
            aug = torch.from_numpy(np.asarray(aug_np, np.float64)).to(dev, non_blocking=True) if augment else None
        else:
            choices = torch.from_numpy(np.stack([np.asarray(d[0]) for d in draws]).astype(np.int32)).to(dev, non_blocking=True)
            augment = draws[0][1] is not None
            aug = torch.from_numpy(np.stack([d[1] for d in draws])).to(dev, non_blocking=True) if augment else None
        row0 = torch.from_numpy(self.row0[sidx]).to(dev, non_blocking=True)
        sidx_d = torch.from_numpy(sidx).to(dev, non_blocking=True)
        fh = self.floor_height[sidx_d].contiguous() if use_height else None
        pc = _ext.prepare_point_clouds(self.verts, row0, choices, multiview=self.multiview if use_multiview else None,
                                       floor_height=fh, aug=aug, mean_rgb=MEAN_COLOR_RGB, use_color=use_color,
                                       use_normal=use_normal)
        boxes = self.boxes[sidx_d].contiguous()
        if augment:
            boxes = _ext.augment_boxes(boxes, aug)
        nb = torch.from_numpy(self.num_bbox[sidx]).to(dev, non_blocking=True)
        out = {"point_clouds": pc, "choices": choices, "target_bboxes": boxes,
               "center_label": boxes[:, :, 0:3].to(torch.float32), "num_bbox": nb,
               "box_label_mask": (torch.arange(MAX_NUM_OBJ, device=dev)[None, :] < nb[:, None]).to(torch.float32)}
        if want_votes:
            votes, mask, overflow = _ext.vote_labels(pc, self.instance_labels, self.semantic_labels, row0, choices,
                                                     self.max_instances, sem_mask_of())
            out["vote_label"], out["vote_label_mask"], out["_instance_overflow"] = votes, mask, overflow
        return out
