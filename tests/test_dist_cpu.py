"""world_size-2 gloo tests (CPU) of the multi-GPU plumbing: scene sharding with no data-path
collective for inference, max-over-ranks timing, and the single flat gradient all-reduce that
replaces DataParallel's reduce in training.  The ops run on the CPU oracle provider
here (tests may use the oracle); the same code drives NCCL ranks on the GPUs."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _small_model():
    from spacap3d_b200.pointnet2_modules import PointnetFPModule, PointnetSAModuleVotes
    torch.manual_seed(0)
    sa1 = PointnetSAModuleVotes(npoint=64, radius=0.6, nsample=8, mlp=[3, 16, 16], use_xyz=True,
                                normalize_xyz=True, bn=False)
    sa2 = PointnetSAModuleVotes(npoint=16, radius=1.2, nsample=8, mlp=[16, 16, 16], use_xyz=True,
                                normalize_xyz=True, bn=False)
    fp = PointnetFPModule(mlp=[16 + 16, 16], bn=False)
    return torch.nn.ModuleList([sa1, sa2, fp])


def _loss(model, xyz, feats):
    sa1, sa2, fp = model
    x1, f1, _ = sa1(xyz, feats)
    x2, f2, _ = sa2(x1, f1)
    up = fp(x1, x2, f1, f2)
    return (up ** 2).sum() / xyz.shape[0]


def _scenes(total):
    from spacap3d_b200.scenes import make_scene_xyz
    xyz = torch.from_numpy(np.stack([make_scene_xyz(70 + i, 400) for i in range(total)], 0))
    g = torch.Generator().manual_seed(1)
    return xyz, torch.randn(total, 3, 400, generator=g)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    import bench
    from spacap3d_b200 import dist as sdist
    total = 6
    xyz, feats = _scenes(total)
    b, e = sdist.shard_range(total, rank, world)
    model = _small_model()
    with bench.swapped_ops(bench.OracleOps(), host_decode=False):
        # inference: each rank runs its shard, no collective; gather only to check
        with torch.no_grad():
            out = _loss(model, xyz[b:e], feats[b:e]) * (e - b)
        # training step: local backward on the shard, then ONE flat all-reduce
        loss = _loss(model, xyz[b:e], feats[b:e]) * (e - b) / total
        loss.backward()
    n = sdist.allreduce_gradients(model, average=False)
    t = sdist.max_over_ranks(1.0 + rank)
    grads = [p.grad.clone() for p in model.parameters()]
    q.put((rank, float(out), n, t, [g.numpy() for g in grads]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_gradient_allreduce():
    import bench
    port = 29500 + (os.getpid() % 2000)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=240) for _ in procs])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process reference on the full batch
    xyz, feats = _scenes(6)
    model = _small_model()
    with bench.swapped_ops(bench.OracleOps(), host_decode=False):
        with torch.no_grad():
            full = float(_loss(model, xyz, feats) * 6)
        _loss(model, xyz, feats).backward()
    want = [p.grad.numpy() for p in model.parameters()]
    assert abs(res[0][1] + res[1][1] - full) <= 1e-4 * abs(full)          # shards partition the batch
    assert res[0][2] == res[1][2] == sum(p.numel() for p in model.parameters())   # one flat bucket
    assert res[0][3] == res[1][3] == 2.0                                     # max over ranks
    for g0, g1, w in zip(res[0][4], res[1][4], want):
        np.testing.assert_allclose(g0, g1, rtol=0, atol=0)                   # ranks agree after the all-reduce
        np.testing.assert_allclose(g0, w, rtol=2e-4, atol=1e-5 * max(1.0, np.abs(w).max()))


def test_shard_range_partitions():
    from spacap3d_b200.dist import shard_range
    for total in (1, 7, 8, 32, 64):
        for world in (1, 2, 4, 8):
            ends = [shard_range(total, r, world) for r in range(world)]
            assert ends[0][0] == 0 and ends[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(ends, ends[1:]))
            sizes = [e - b for b, e in ends]
            assert max(sizes) - min(sizes) <= 1
