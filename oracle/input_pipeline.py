"""TEST ORACLE (not product code): numpy restatement of the point-data part of the reference's
`ScannetReferenceDataset.__getitem__` (lib/dataset.py:291-531), split at the same places as the C ABI of
csrc/input_pipeline.cu.  Only tests/, __graft_entry__.smoke() and the cpu_baseline legs may import it.

Pinned: tests/golden/input_ref.npz was produced by oracle/make_golden_input.py, which executes the reference's own
function bodies (random_sampling, rotx/roty/rotz, rotate_aligned_boxes_along_axis, _translate -- extracted from the
reference sources by AST because their modules import plyfile / trimesh / easydict, which are not installed) in the
reference's statement order; tests/test_input_pipeline.py checks every function here against it.
numpy semantics are those of numpy >= 2.0 (this container: 2.3.5): np.percentile of a float32 column is evaluated in
float32; the cloud stays float32 between augmentation steps.
"""
import numpy as np

MEAN_COLOR_RGB = np.array([109.8, 97.2, 83.8])      # lib/dataset.py:28
NYU40IDS = np.array([3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 23, 24, 25, 26, 27, 28, 29,
                     30, 31, 32, 33, 34, 35, 36, 37, 38, 39, 40])   # data/scannet/model_util_scannet.py:88


def floor_height(z, q=0.99):
    """np.percentile(z, q) for float32 z, spelled out (numpy/lib/_function_base_impl.py: percentile -> _quantile ->
    _lerp): quantile, virtual index, gamma and the interpolation are all float32.   lib/dataset.py:331"""
    z = np.asarray(z, np.float32)
    n = z.shape[0]
    quant = np.float32(q) / np.float32(100)
    v = np.float32(n - 1) * quant
    i = int(np.floor(v))
    t = np.float32(v - np.float32(i))
    s = np.sort(z)
    if v >= n - 1:
        i = n - 1
    a, b = s[i], s[min(i + 1, n - 1)]
    d = np.float32(b - a)
    r = np.float32(a + np.float32(d * t))
    if t >= 0.5:
        r = np.float32(b - np.float32(d * np.float32(np.float32(1) - t)))
    return r


def _mats(aug):
    return aug[2:11].reshape(3, 3), aug[11:20].reshape(3, 3), aug[20:29].reshape(3, 3)


def prepare_point_cloud(verts, choices, multiview=None, fh=None, aug=None, use_color=False, use_normal=False):
    """One item: (M,9) f32 vertex table -> (P,C) f32.   lib/dataset.py:309-335, 366-404"""
    cols = [verts[:, 0:3]]
    if use_color:
        cols.append(((verts[:, 3:6] - MEAN_COLOR_RGB) / 256.0).astype(np.float32))     # :314 (float64 -> float32 store)
    if use_normal:
        cols.append(verts[:, 6:9])                                                    # :317-319
    if multiview is not None:
        cols.append(multiview)                                                        # :321-328
    if fh is not None:
        cols.append((verts[:, 2] - np.float32(fh))[:, None])                           # :330-333
    pc = np.concatenate(cols, 1).astype(np.float32)[choices]                           # :335
    if aug is not None:
        if aug[0] != 0:
            pc[:, 0] = -1 * pc[:, 0]                                                  # :369
        if aug[1] != 0:
            pc[:, 1] = -1 * pc[:, 1]                                                  # :379
        for R in _mats(aug):
            pc[:, 0:3] = np.dot(pc[:, 0:3], np.transpose(R))                          # :390,396,402
        coords = pc[:, :3]
        coords += [aug[29], aug[30], aug[31]]                                         # :240
    return pc


def vote_labels(pc, instance_labels, semantic_labels):
    """pc (P,>=3) f32 and the SAMPLED labels -> vote_label (P,9) f32, vote_label_mask (P,) i64.  lib/dataset.py:421-431"""
    P = pc.shape[0]
    votes = np.zeros([P, 3])
    mask = np.zeros(P)
    for i_instance in np.unique(instance_labels):
        ind = np.where(instance_labels == i_instance)[0]
        if semantic_labels[ind[0]] in NYU40IDS:
            x = pc[ind, :3]
            center = 0.5 * (x.min(0) + x.max(0))
            votes[ind, :] = center - x
            mask[ind] = 1.0
    return np.tile(votes, (1, 3)).astype(np.float32), mask.astype(np.int64)


def _rotate_boxes(boxes, R, axis):
    """rotate_aligned_boxes_along_axis, data/scannet/model_util_scannet.py:47-82"""
    centers, lengths = boxes[:, 0:3], boxes[:, 3:6]
    new_centers = np.dot(centers, np.transpose(R))
    i1, i2 = {"x": (1, 2), "y": (0, 2), "z": (0, 1)}[axis]
    d1, d2 = lengths[:, i1] / 2.0, lengths[:, i2] / 2.0
    n1 = np.zeros((d1.shape[0], 4))
    n2 = np.zeros((d1.shape[0], 4))
    for i, (s1, s2) in enumerate([(-1, -1), (1, -1), (1, 1), (-1, 1)]):
        c = np.zeros((d1.shape[0], 3))
        c[:, 0] = s1 * d1
        c[:, 1] = s2 * d2
        c = np.dot(c, np.transpose(R))
        n1[:, i] = c[:, 0]
        n2[:, i] = c[:, 1]
    new = lengths.copy()
    new[:, i1] = 2.0 * np.max(n1, 1)
    new[:, i2] = 2.0 * np.max(n2, 1)
    return np.concatenate([new_centers, new], axis=1)


def augment_boxes(boxes, aug):
    """boxes (K,6) f64 -> (K,6) f64.   lib/dataset.py:369-404"""
    b = np.array(boxes, np.float64)
    if aug[0] != 0:
        b[:, 0] = -1 * b[:, 0]
    if aug[1] != 0:
        b[:, 1] = -1 * b[:, 1]
    for R, axis in zip(_mats(aug), "xyz"):
        b = _rotate_boxes(b, R, axis)
    b[:, :3] += [aug[29], aug[30], aug[31]]
    return b
