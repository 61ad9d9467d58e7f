"""Seeded synthetic 'pre-processed ScanNet scenes' for the input-pipeline tests (shared with
oracle/make_golden_input.py).  Layout of data/scannet/batch_load_scannet_data.py:75-80: mesh_vertices (M,9) f32 =
xyz, rgb (0..255), normal; instance / semantic labels per vertex; instance_bboxes (K,8) = centre, lengths, nyu40 id,
object id."""
import numpy as np

SEM_POOL = np.array([1, 2, 22, 3, 4, 5, 7, 8, 14, 24, 33, 39, 40])     # walls/floor/ceiling (excluded) + object classes

# name -> (M, P, n_instances, use_color, use_normal, use_multiview, use_height, augment, seed)
CASES = {
    "xyz_h_aug": (12000, 8192, 40, False, False, False, True, True, 11),
    "color_normal_h_aug": (9000, 4096, 25, True, True, False, True, True, 12),
    "multiview_h_aug": (3000, 2000, 12, False, True, True, True, True, 13),
    "replace_eval": (1500, 2000, 6, True, False, False, True, False, 14),
    "noheight_aug": (5000, 4000, 300, False, False, False, False, True, 15),
}


def make_scene(M, n_inst, seed, multiview=False):
    rng = np.random.default_rng(seed)
    v = np.zeros((M, 9), np.float32)
    v[:, 0:3] = (rng.random((M, 3)) * [8.0, 6.0, 3.0] - [4.0, 3.0, 0.05]).astype(np.float32)
    v[rng.random(M) < 0.3, 2] = np.float32(0.0)                       # a flat floor: duplicated z values
    v[:, 3:6] = rng.integers(0, 256, (M, 3)).astype(np.float32)
    nrm = rng.standard_normal((M, 3))
    v[:, 6:9] = (nrm / np.linalg.norm(nrm, axis=1, keepdims=True)).astype(np.float32)
    # instances: 0 = unannotated (35 %), others = nearest of n_inst random seeds
    seeds = rng.random((n_inst, 3)) * [8.0, 6.0, 3.0] - [4.0, 3.0, 0.0]
    near = np.argmin(((v[:, None, 0:3] - seeds[None]) ** 2).sum(-1), 1) if M * n_inst < 5e6 else \
        rng.integers(0, n_inst, M)
    inst = np.where(rng.random(M) < 0.35, 0, near + 1).astype(np.uint32)
    sem_of = SEM_POOL[rng.integers(0, len(SEM_POOL), n_inst + 1)]
    sem = sem_of[inst].astype(np.uint32)
    flip = rng.random(M) < 0.1                                         # inconsistent labels inside an instance:
    sem[flip] = SEM_POOL[rng.integers(0, len(SEM_POOL), flip.sum())]   # "label of the first sampled point" matters
    K = min(n_inst, 140)                                               # > MAX_NUM_OBJ exercises the truncation
    bboxes = np.zeros((K, 8), np.float64)
    bboxes[:, 0:3] = seeds[:K]
    bboxes[:, 3:6] = rng.uniform(0.2, 2.5, (K, 3))
    bboxes[:, 6] = sem_of[1:K + 1]
    bboxes[:, 7] = np.arange(K)
    mv = None
    if multiview:
        mv = np.maximum(rng.standard_normal((M, 128)), 0).astype(np.float32)
        mv[rng.random(M) < 0.2] = 0
    return v, inst, sem, bboxes, mv


def case(name):
    M, P, n_inst, col, nrm, mv, h, aug, seed = CASES[name]
    v, inst, sem, bb, mvf = make_scene(M, n_inst, seed, mv)
    return dict(verts=v, inst=inst, sem=sem, bboxes=bb, multiview=mvf, P=P, use_color=col, use_normal=nrm,
                use_multiview=mv, use_height=h, augment=aug, seed=seed)


def percentile_inputs():
    """float32 columns of many sizes / duplicate patterns for the np.percentile(z, 0.99) restatement."""
    rng = np.random.default_rng(0)
    out = []
    for i in range(40):
        n = int(rng.integers(2, 70000))
        z = (rng.standard_normal(n) * rng.uniform(0.1, 3) + rng.uniform(-2, 2)).astype(np.float32)
        if i % 3 == 0:
            z = np.round(z, 1)
        out.append(z)
    out.append(np.float32([1.5]))
    out.append(np.float32([2.0, -1.0]))
    out.append(np.zeros(50000, np.float32))
    return out
