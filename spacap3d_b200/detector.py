"""The callers of the hot path: VoteNet-style detector (PointNet++ backbone -> Hough voting ->
vote aggregation + proposal head) with the reference's module tree and state-dict keys
(`backbone_net.sa1.mlp_module.layer0.conv.weight`, `vgen.conv1.weight`,
`proposal.vote_aggregation...`, `proposal.proposal.0.weight`), so the reference's
`pretrained/*/model.pth` detector checkpoints load with strict=False exactly as
scripts/train.py:170-181 does.

Follows models/backbone_module.py:23-129, models/voting_module.py:13-61,
models/proposal_module.py:20-158 and the detection branch of models/SpaCapNet.py:47-74.
SURVEY row N1: the reference's `decode_pred_box` (proposal_module.py:81-104) round-trips through
host numpy with a Python loop over the batch on every forward; here the box corners are decoded
on the device (float64, same arithmetic: ScanNet boxes are axis aligned, heading is always 0,
data/scannet/model_util_scannet.py:130-140) so the forward has no host synchronisation.
The captioner (models/transformer_captioner.py) is out of scope.
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _ext
from ._ext import HALF
from .pointnet2_modules import (INLINE_MAX_FEATURES, FoldedChain, PointnetFPModule, PointnetSAModuleVotes, _cache_of,
                                attach_pm, fast_eval_ok, get_pm, get_pm_pair)

# ScanNet per-class mean box sizes (18 x 3, float64, every digit: the decoded corners must equal the reference's
# bit for bit): dataset metadata shipped by the reference as
# data/scannet/meta_data/scannet_reference_means.npz, read by ScannetDatasetConfig
# (data/scannet/model_util_scannet.py:90).
SCANNET_MEAN_SIZE_ARR = np.array([
    [0.7750491029714929, 0.9489772784305719, 0.9654205889420883],
    [1.8690326739217817, 1.8321471223511647, 1.1922299150646347],
    [0.6121477783923587, 0.6192873075057846, 0.7048084833710475],
    [1.4411389838393118, 1.6045203579823017, 0.8365229505964112],
    [1.0478072557954565, 1.2016418836390361, 0.6345700676484581],
    [0.5610123179013166, 0.6084721692226233, 1.7195040055943263],
    [1.0789489470730143, 0.8203399609681988, 1.1692119917347412],
    [0.8417109198057999, 1.3504794475570598, 1.689892503247653],
    [0.2305173710207977, 0.4764049876932717, 0.5656925618884787],
    [1.4548489887322953, 1.9711989456815506, 0.28643280467880305],
    [1.0785803060791836, 1.5370511310202535, 0.8650190604735265],
    [1.4311964378217468, 0.7692311116413818, 1.6498267253793382],
    [0.6296919388045009, 0.7087128690976665, 1.314335867333314],
    [0.4392503422374527, 0.41569593879911637, 1.7000274790657892],
    [0.5850446242623347, 0.5787843832293073, 0.7202961145680844],
    [0.5115869258698381, 0.5096067340403306, 0.3128736034402105],
    [1.1732075942887201, 1.0598714035004377, 0.5181252788752317],
    [0.43294385021345605, 0.5193350711870748, 0.4843745602902239],
], dtype=np.float64)

NUM_CLASS = 18
NUM_HEADING_BIN = 1
NUM_SIZE_CLUSTER = 18

# corner sign pattern of utils/box_util.py:360-383 (x: l, y: w, z: h)
_CORNER_SIGNS = torch.tensor([[1, 1, 1], [1, -1, 1], [-1, -1, 1], [-1, 1, 1],
                              [1, 1, -1], [1, -1, -1], [-1, -1, -1], [-1, 1, -1]],
                             dtype=torch.float64)


class Pointnet2Backbone(nn.Module):
    """SA1-4 (2048/1024/512/256 centres) + FP1-2; input (B,N,3+C) with C = input_feature_dim."""

    SA_SPECS = ((2048, 0.2, 64, (64, 64, 128)), (1024, 0.4, 32, (128, 128, 256)),
                (512, 0.8, 16, (128, 128, 256)), (256, 1.2, 16, (128, 128, 256)))

    def __init__(self, input_feature_dim=0):
        super().__init__()
        self.input_feature_dim = input_feature_dim
        cin = input_feature_dim
        for i, (npoint, radius, nsample, widths) in enumerate(self.SA_SPECS, 1):
            setattr(self, "sa%d" % i, PointnetSAModuleVotes(
                npoint=npoint, radius=radius, nsample=nsample, mlp=[cin] + list(widths),
                use_xyz=True, normalize_xyz=True))
            cin = widths[-1]
        self.fp1 = PointnetFPModule(mlp=[256 + 256, 256, 256])
        self.fp2 = PointnetFPModule(mlp=[256 + 256, 256, 256])

    @staticmethod
    def _break_up_pc(pc):
        xyz = pc[..., :3].contiguous()
        if pc.size(-1) <= 3:
            return xyz, None
        C = pc.size(-1) - 3
        if C > INLINE_MAX_FEATURES and fast_eval_ok(pc) and pc.dtype == torch.float32:
            # many input channels (multiview features): the fused SA1 wants them point-major in fp16 -- which is
            # the layout the cloud already has.  One pass makes that copy (zero-padded to a multiple of 8 channels
            # for TMA); the channel-major tensor of the reference API stays a view and is never materialised.
            Kp = (C + 7) // 8 * 8
            pm = torch.zeros((pc.shape[0], pc.shape[1], Kp), dtype=HALF, device=pc.device) if Kp != C else \
                torch.empty((pc.shape[0], pc.shape[1], Kp), dtype=HALF, device=pc.device)
            pm[..., :C].copy_(pc[..., 3:])
            return xyz, attach_pm(pc[..., 3:].transpose(1, 2), pm)
        return xyz, pc[..., 3:].transpose(1, 2).contiguous()

    def forward(self, data_dict):
        xyz, features = self._break_up_pc(data_dict["point_clouds"])
        xyz, features, inds = self.sa1(xyz, features)
        data_dict["sa1_inds"], data_dict["sa1_xyz"], data_dict["sa1_features"] = inds, xyz, features
        xyz, features, inds = self.sa2(xyz, features)
        data_dict["sa2_inds"], data_dict["sa2_xyz"], data_dict["sa2_features"] = inds, xyz, features
        xyz, features, _ = self.sa3(xyz, features)
        data_dict["sa3_xyz"], data_dict["sa3_features"] = xyz, features
        xyz, features, _ = self.sa4(xyz, features)
        data_dict["sa4_xyz"], data_dict["sa4_features"] = xyz, features
        features = self.fp1(data_dict["sa3_xyz"], data_dict["sa4_xyz"], data_dict["sa3_features"],
                            data_dict["sa4_features"])
        features = self.fp2(data_dict["sa2_xyz"], data_dict["sa3_xyz"], data_dict["sa2_features"],
                            features)
        data_dict["fp2_features"] = features
        data_dict["fp2_xyz"] = data_dict["sa2_xyz"]
        num_seed = data_dict["fp2_xyz"].shape[1]
        # relies on FPS over an FPS-ordered prefix returning 0..n-1, like the reference
        data_dict["fp2_inds"] = data_dict["sa1_inds"][:, 0:num_seed]
        return data_dict


class VotingModule(nn.Module):
    def __init__(self, vote_factor, seed_feature_dim):
        super().__init__()
        self.vote_factor = vote_factor
        self.in_dim = self.out_dim = seed_feature_dim
        self.conv1 = nn.Conv1d(self.in_dim, self.in_dim, 1)
        self.conv2 = nn.Conv1d(self.in_dim, self.in_dim, 1)
        self.conv3 = nn.Conv1d(self.in_dim, (3 + self.out_dim) * self.vote_factor, 1)
        self.bn1 = nn.BatchNorm1d(self.in_dim)
        self.bn2 = nn.BatchNorm1d(self.in_dim)

    def forward(self, seed_xyz, seed_features):
        """seed_xyz (B,S,3), seed_features (B,D,S) -> vote_xyz (B,S*vf,3), vote_features (B,D,S*vf)"""
        B, S = seed_xyz.shape[0], seed_xyz.shape[1]
        net = F.relu(self.bn1(self.conv1(seed_features)))
        net = F.relu(self.bn2(self.conv2(net)))
        net = self.conv3(net).transpose(2, 1).view(B, S, self.vote_factor, 3 + self.out_dim)
        vote_xyz = (seed_xyz.unsqueeze(2) + net[..., 0:3]).contiguous().view(B, S * self.vote_factor, 3)
        vote_features = seed_features.transpose(2, 1).unsqueeze(2) + net[..., 3:]
        vote_features = vote_features.contiguous().view(B, S * self.vote_factor, self.out_dim)
        return vote_xyz, vote_features.transpose(2, 1).contiguous()

    def forward_normalized_fast(self, seed_xyz, seed_features):
        """Eval fast path returning (vote_xyz, L2-normalised vote_features) or None: three tcgen05 launches
        (spc_pm_linear: BN folded, fp16-pair operands, fp32-grade); offsets, residual, normalisation and both output
        layouts are the epilogue of the last one."""
        pm, pm_lo = get_pm_pair(seed_features)
        if (self.training or self.vote_factor != 1 or pm is None or self.out_dim > 256 or self.in_dim % 64
                or not fast_eval_ok(seed_xyz, seed_features)):
            return None
        chain = _cache_of(self, FoldedChain).get([(self.conv1, self.bn1), (self.conv2, self.bn2), (self.conv3, None)],
                                                 last_rows_rotate=3)          # features first, xyz offsets last
        B, S, D = pm.shape
        h = pm.reshape(B * S, D) if pm_lo is None else (pm.reshape(B * S, D), pm_lo.reshape(B * S, D))
        h = _ext.pm_linear(h, chain[0][0], chain[0][1], _ext.PM_HIDDEN, S)
        h = _ext.pm_linear(h, chain[1][0], chain[1][1], _ext.PM_HIDDEN, S)
        vote_xyz, vote_features, (hi, lo) = _ext.pm_linear(h, chain[2][0], chain[2][1], _ext.PM_VOTE, S,
                                                           seed_cm=seed_features.contiguous(),
                                                           seed_xyz=seed_xyz.contiguous())
        return vote_xyz, attach_pm(vote_features, hi.view(B, S, D), lo.view(B, S, D))


class ProposalModule(nn.Module):
    def __init__(self, num_class, num_heading_bin, num_size_cluster, mean_size_arr, num_proposal,
                 sampling="vote_fps", seed_feat_dim=256, size_decoded=False):
        super().__init__()
        self.num_class = num_class
        self.num_heading_bin = num_heading_bin
        self.num_size_cluster = num_size_cluster
        self.mean_size_arr = mean_size_arr
        self.num_proposal = num_proposal
        self.sampling = sampling
        self.seed_feat_dim = seed_feat_dim
        self.size_decoded = size_decoded
        self.vote_aggregation = PointnetSAModuleVotes(
            npoint=num_proposal, radius=0.3, nsample=16, mlp=[seed_feat_dim, 128, 128, 128],
            use_xyz=True, normalize_xyz=True)
        out_dim = 2 + 3 + num_heading_bin * 2 + num_size_cluster * 4 + num_class
        self.proposal = nn.Sequential(
            nn.Conv1d(128, 128, 1, bias=False), nn.BatchNorm1d(128), nn.ReLU(),
            nn.Conv1d(128, 128, 1, bias=False), nn.BatchNorm1d(128), nn.ReLU(),
            nn.Conv1d(128, out_dim, 1))
        # constants used by the decode, kept as non-persistent buffers so they follow .to(device)
        # without entering the state dict (the reference rebuilds them from numpy every forward)
        self.register_buffer("_mean_size_f32", torch.from_numpy(mean_size_arr.astype(np.float32)),
                             persistent=False)
        self.register_buffer("_mean_size_f64", torch.from_numpy(mean_size_arr.astype(np.float64)),
                             persistent=False)
        self.register_buffer("_corner_signs", _CORNER_SIGNS.clone(), persistent=False)

    def forward(self, xyz, features, data_dict):
        xyz, features, fps_inds = self.vote_aggregation(xyz, features)
        data_dict["aggregated_vote_xyz"] = xyz
        data_dict["aggregated_vote_features"] = features.permute(0, 2, 1).contiguous()
        data_dict["aggregated_vote_inds"] = fps_inds
        pm = get_pm(features)
        if not self.training and pm is not None and pm.shape[2] % 64 == 0 and fast_eval_ok(features):
            chain = _cache_of(self, FoldedChain).get(
                [(self.proposal[0], self.proposal[1]), (self.proposal[3], self.proposal[4]), (self.proposal[6], None)])
            B, K, D = pm.shape
            h = _ext.pm_linear(pm.reshape(B * K, D), chain[0][0], chain[0][1], _ext.PM_HIDDEN, K)
            h = _ext.pm_linear(h, chain[1][0], chain[1][1], _ext.PM_HIDDEN, K)
            t = _ext.pm_linear(h, chain[2][0], chain[2][1], _ext.PM_OUT_PM32, K).view(B, K, -1)   # fp32 head output
            return self.decode_scores(None, data_dict, net_transposed=t)
        net = self.proposal(features)
        return self.decode_scores(net, data_dict)

    def decode_pred_box(self, data_dict):
        """(B,K,8,3) float64 box corners, on the device (reference: host numpy round trip)."""
        center = data_dict["center"].detach().double()                                # (B,K,3)
        size_class = torch.argmax(data_dict["size_scores"], -1)                       # (B,K)
        gather_idx = size_class[..., None, None].expand(-1, -1, 1, 3)
        residual = torch.gather(data_dict["size_residuals"], 2, gather_idx).squeeze(2).detach()
        box_size = self._mean_size_f64[size_class] + residual.double()                # class2size
        # heading is identically zero for ScanNet => R = I and the matmul is exact
        corners = self._corner_signs * (box_size / 2).unsqueeze(-2)                   # (B,K,8,3)
        return corners + center.unsqueeze(-2)

    def decode_scores(self, net, data_dict, net_transposed=None):
        nh, ns = self.num_heading_bin, self.num_size_cluster
        t = net.transpose(2, 1).contiguous() if net_transposed is None else net_transposed   # (B,K,out)
        B, K = t.shape[0], t.shape[1]
        objectness_scores = t[:, :, 0:2]
        center = data_dict["aggregated_vote_xyz"] + t[:, :, 2:5]
        heading_scores = t[:, :, 5:5 + nh]
        heading_residuals_normalized = t[:, :, 5 + nh:5 + nh * 2]
        size_scores = t[:, :, 5 + nh * 2:5 + nh * 2 + ns]
        size_residuals_normalized = t[:, :, 5 + nh * 2 + ns:5 + nh * 2 + ns * 4].view(B, K, ns, 3)
        sem_cls_scores = t[:, :, 5 + nh * 2 + ns * 4:]
        mean = self._mean_size_f32.unsqueeze(0).unsqueeze(0)
        data_dict["objectness_scores"] = objectness_scores
        data_dict["center"] = center
        data_dict["heading_scores"] = heading_scores
        data_dict["heading_residuals_normalized"] = heading_residuals_normalized
        data_dict["heading_residuals"] = heading_residuals_normalized * (np.pi / nh)
        data_dict["size_scores"] = size_scores
        data_dict["size_residuals_normalized"] = size_residuals_normalized
        data_dict["size_residuals"] = size_residuals_normalized * mean
        if self.size_decoded:
            size_recover = data_dict["size_residuals"] + mean
            cls = torch.argmax(size_scores, -1)[..., None, None].repeat(1, 1, 1, 3)
            data_dict["pred_size"] = torch.gather(size_recover, 2, cls).squeeze(2)
        data_dict["sem_cls_scores"] = sem_cls_scores
        data_dict["bbox_corner"] = self.decode_pred_box(data_dict)
        data_dict["bbox_feature"] = data_dict["aggregated_vote_features"]
        data_dict["bbox_mask"] = objectness_scores.argmax(-1)
        data_dict["bbox_sems"] = sem_cls_scores.argmax(-1)
        data_dict["sem_cls"] = sem_cls_scores.argmax(-1)
        return data_dict


class VoteNetDetector(nn.Module):
    """Detection branch of SpaCapNet (models/SpaCapNet.py:47-74) with the same child names."""

    def __init__(self, input_feature_dim=0, num_proposal=256, vote_factor=1, num_class=NUM_CLASS,
                 num_heading_bin=NUM_HEADING_BIN, num_size_cluster=NUM_SIZE_CLUSTER,
                 mean_size_arr=SCANNET_MEAN_SIZE_ARR, sampling="vote_fps"):
        super().__init__()
        assert mean_size_arr.shape[0] == num_size_cluster
        self.input_feature_dim = input_feature_dim
        self.backbone_net = Pointnet2Backbone(input_feature_dim=input_feature_dim)
        self.vgen = VotingModule(vote_factor, 256)
        self.proposal = ProposalModule(num_class, num_heading_bin, num_size_cluster, mean_size_arr,
                                       num_proposal, sampling)

    def forward(self, data_dict):
        data_dict = self.backbone_net(data_dict)
        xyz, features = data_dict["fp2_xyz"], data_dict["fp2_features"]
        data_dict["seed_inds"] = data_dict["fp2_inds"]
        data_dict["seed_xyz"] = xyz
        data_dict["seed_features"] = features
        fast = self.vgen.forward_normalized_fast(xyz, features)
        if fast is not None:
            xyz, features = fast
        else:
            xyz, features = self.vgen(xyz, features)
            features = features.div(torch.norm(features, p=2, dim=1).unsqueeze(1))
        data_dict["vote_xyz"] = xyz
        data_dict["vote_features"] = features
        return self.proposal(xyz, features, data_dict)
