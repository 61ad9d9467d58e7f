"""Eval fast paths of the callers (FP modules, voting, proposal head, whole detector) against the
reference op sequence in fp32 (FAST_PATHS off, TF32 off).  The fast paths run the 1x1-conv chains
as fp16 GEMMs (fp32 accumulate) => tolerance 1e-2 .. 3e-2 of the tensor's max magnitude (north_star:
"1e-2 for fp16 MLP"; errors compound over the nine fp16 stages of the full detector)."""
import numpy as np
import pytest
import torch

from spacap3d_b200._ext import HALF
from spacap3d_b200.pointnet2_modules import attach_pm, get_pm

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


class fp32_reference:
    def __enter__(self):
        from spacap3d_b200 import pointnet2_modules as M
        self.M = M
        self.old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, M.FAST_PATHS)
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        M.FAST_PATHS = False

    def __exit__(self, *a):
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, self.M.FAST_PATHS = self.old


def _randomize_bn(model, seed):
    g = torch.Generator(device="cpu").manual_seed(seed)
    for mod in model.modules():
        if isinstance(mod, (torch.nn.BatchNorm1d, torch.nn.BatchNorm2d)):
            mod.running_mean.copy_(torch.randn(mod.num_features, generator=g) * 0.1)
            mod.running_var.copy_(torch.rand(mod.num_features, generator=g) * 0.5 + 0.75)
            mod.weight.data.copy_(torch.rand(mod.num_features, generator=g) * 0.5 + 0.75)
            mod.bias.data.copy_(torch.randn(mod.num_features, generator=g) * 0.05)


def _rel(a, b):
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()


def test_three_nn_weights_matches_module_formula():
    from spacap3d_b200 import _ext
    g = torch.Generator(device="cpu").manual_seed(0)
    unknown = (torch.rand(3, 700, 3, generator=g) * 4 - 2).to(DEV)
    known = (torch.rand(3, 300, 3, generator=g) * 4 - 2).to(DEV)
    known[0, :50] = unknown[0, :50]                                 # exact zero distances
    idx, w = _ext.three_nn_weights(unknown, known)
    d2, ridx = _ext.three_nn(unknown, known)
    assert torch.equal(idx, ridx)
    dist = torch.sqrt(d2)
    recip = 1.0 / (dist + 1e-8)
    want = recip / recip.sum(dim=2, keepdim=True)
    torch.testing.assert_close(w, want, rtol=1e-6, atol=1e-7)


def test_fp_module_fast_path():
    from spacap3d_b200.pointnet2_modules import PointnetFPModule
    torch.manual_seed(1)
    fp = PointnetFPModule(mlp=[256 + 256, 256, 256]).to(DEV).eval()
    _randomize_bn(fp, 2)
    g = torch.Generator(device="cpu").manual_seed(3)
    unknown = (torch.rand(2, 512, 3, generator=g) * 6 - 3).to(DEV)
    known = unknown[:, ::2].contiguous()
    uf = torch.relu(torch.randn(2, 256, 512, generator=g)).to(DEV)
    kf = torch.relu(torch.randn(2, 256, 256, generator=g)).to(DEV)
    attach_pm(uf, uf.transpose(1, 2).contiguous().to(HALF))
    attach_pm(kf, kf.transpose(1, 2).contiguous().to(HALF))
    with torch.no_grad():
        fast = fp(unknown, known, uf, kf)
        assert get_pm(fast) is not None                             # the fast path ran
        with fp32_reference():
            ref = fp(unknown, known, uf, kf)
    assert fast.shape == ref.shape == (2, 256, 512)
    assert _rel(fast, ref) <= 1e-2
    assert _rel(get_pm(fast).float().transpose(1, 2), ref) <= 1e-2


def test_voting_and_proposal_fast_paths():
    from spacap3d_b200.detector import ProposalModule, VotingModule, SCANNET_MEAN_SIZE_ARR
    torch.manual_seed(4)
    vgen = VotingModule(1, 256).to(DEV).eval()
    prop = ProposalModule(18, 1, 18, SCANNET_MEAN_SIZE_ARR, 256).to(DEV).eval()
    _randomize_bn(vgen, 5)
    _randomize_bn(prop, 6)
    g = torch.Generator(device="cpu").manual_seed(7)
    seed_xyz = (torch.rand(2, 1024, 3, generator=g) * 6 - 3).to(DEV)
    sf = torch.relu(torch.randn(2, 256, 1024, generator=g)).to(DEV)
    attach_pm(sf, sf.transpose(1, 2).contiguous().to(HALF))
    with torch.no_grad():
        fast = vgen.forward_normalized_fast(seed_xyz, sf)
        assert fast is not None
        vx, vf = fast
        rx, rf = vgen(seed_xyz, sf)
        rf = rf.div(torch.norm(rf, p=2, dim=1).unsqueeze(1))
        assert (vx - rx).abs().max().item() <= 1e-2 * (rx - seed_xyz).abs().max().item() + 1e-5   # offsets
        assert _rel(vf, rf) <= 1e-2
        assert _rel(get_pm(vf).float().transpose(1, 2), rf) <= 1e-2
        # proposal module on identical inputs: FPS indices must agree bit for bit, scores to 2e-2
        d_fast = prop(vx, vf, {})
        with fp32_reference():
            vf_plain = vf.clone()
            d_ref = prop(vx, vf_plain, {})
    assert torch.equal(d_fast["aggregated_vote_inds"], d_ref["aggregated_vote_inds"])
    for k in ("objectness_scores", "center", "size_scores", "sem_cls_scores", "size_residuals"):
        assert d_fast[k].shape == d_ref[k].shape
        assert _rel(d_fast[k], d_ref[k]) <= 2e-2, k
    assert d_fast["bbox_corner"].dtype == torch.float64 and d_fast["bbox_corner"].shape == (2, 256, 8, 3)


def test_detector_backbone_and_votes_fast_vs_fp32():
    """Whole backbone + voting: indices of SA1-4 bit-exact (they do not depend on features), seed and
    vote features within 3e-2 of max after eight fp16 stages."""
    from spacap3d_b200.detector import VoteNetDetector
    from spacap3d_b200.scenes import make_scene
    torch.manual_seed(0)
    model = VoteNetDetector(input_feature_dim=1).to(DEV).eval()
    _randomize_bn(model, 11)
    pc = torch.from_numpy(np.stack([make_scene(31, 20000), make_scene(32, 20000, with_replacement=True)], 0)).to(DEV)
    with torch.no_grad():
        fast = model({"point_clouds": pc})
        with fp32_reference():
            ref = model({"point_clouds": pc})
    for k in ("sa1_inds", "sa2_inds", "sa1_xyz", "sa4_xyz", "seed_inds"):
        assert torch.equal(fast[k], ref[k]), k
    for k, tol in (("sa1_features", 1e-2), ("sa4_features", 2e-2), ("fp2_features", 3e-2), ("vote_features", 3e-2)):
        assert _rel(fast[k], ref[k]) <= tol, (k, _rel(fast[k], ref[k]))
    off_f, off_r = fast["vote_xyz"] - fast["seed_xyz"], ref["vote_xyz"] - ref["seed_xyz"]
    assert (off_f - off_r).abs().max().item() <= 3e-2 * off_r.abs().max().item()
    for k in ("objectness_scores", "center", "bbox_corner", "sem_cls_scores", "aggregated_vote_inds"):
        assert fast[k].shape == ref[k].shape and torch.isfinite(fast[k].double()).all()


def test_graphed_pipeline_equals_eager():
    """CUDA-graph replay on several streams (incl. the culled FPS kernel it switches on) returns exactly
    what the eager forward returns, batch after batch."""
    from spacap3d_b200 import _lib
    from spacap3d_b200.detector import VoteNetDetector
    from spacap3d_b200.pipeline import GraphedDetector
    from spacap3d_b200.scenes import make_scene
    torch.manual_seed(0)
    model = VoteNetDetector(input_feature_dim=1).to(DEV).eval()
    _randomize_bn(model, 5)
    batches = [torch.from_numpy(np.stack([make_scene(200 + 2 * i, 20000), make_scene(201 + 2 * i, 20000)], 0)).to(DEV)
               for i in range(4)]
    keys = ("sa1_inds", "seed_inds", "vote_xyz", "aggregated_vote_inds", "objectness_scores", "center",
            "sem_cls_scores", "bbox_corner")
    with torch.no_grad():
        want = [{k: model({"point_clouds": b})[k].clone() for k in keys} for b in batches]
    runner = GraphedDetector(model, batches[0], n_streams=3, result_keys=keys)
    for rep in range(2):
        slots = [runner.submit(b) for b in batches[:3]]
        for s, w in zip(slots, want[:3]):
            runner.wait(s)
            for k in keys:
                assert torch.equal(runner.outputs[s][k], w[k]), (rep, k)
        s = runner.submit(batches[3], to_host=True)
        with pytest.raises(RuntimeError):                 # the slot's host results have not been collected yet
            for _ in range(3):
                runner.submit(batches[3], to_host=True)
        runner.close()
        host = runner.host_out[s]
        for k in keys:
            assert torch.equal(host[k], want[3][k].cpu()), k


@pytest.mark.parametrize("name,kw,dim", [
    ("xyz_only", dict(use_height=False), 0),
    ("rgb_normal_height", dict(use_color=True, use_normal=True), 7),            # BASELINE configs[2]
    ("multiview_normal_height", dict(use_multiview=True, use_normal=True), 132),  # BASELINE configs[3]
])
def test_detector_other_feature_configs(name, kw, dim):
    """The other input layouts of BASELINE.json (parity cases, not bench lines): sampling / grouping indices
    bit-exact, features within the fp16 tolerance of the fp32 reference op sequence."""
    from spacap3d_b200.detector import VoteNetDetector
    from spacap3d_b200.scenes import make_scene
    torch.manual_seed(1)
    model = VoteNetDetector(input_feature_dim=dim).to(DEV).eval()
    _randomize_bn(model, 21)
    pc = torch.from_numpy(np.stack([make_scene(51, 20000, **kw), make_scene(52, 20000, **kw)], 0)).to(DEV)
    assert pc.shape[2] == 3 + dim
    with torch.no_grad():
        fast = model({"point_clouds": pc})
        with fp32_reference():
            ref = model({"point_clouds": pc})
    for k in ("sa1_inds", "sa2_inds", "sa1_xyz", "sa4_xyz", "seed_inds"):
        assert torch.equal(fast[k], ref[k]), k
    for k, tol in (("sa1_features", 1e-2), ("sa4_features", 2e-2), ("fp2_features", 3e-2), ("vote_features", 3e-2)):
        assert _rel(fast[k], ref[k]) <= tol, (k, _rel(fast[k], ref[k]))
