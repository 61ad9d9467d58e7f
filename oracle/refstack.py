"""Import the staged, unmodified reference detector stack (baseline/_ref/, see stage_reference.py)
under one of two bindings of its extension boundary.

TEST INFRASTRUCTURE ONLY: tests/ and bench.py --impl reference.

  load_stack("reference")  stack A: the reference's own lib/pointnet2/*.py + models/*.py, with
                           `pointnet2._ext` = the reference's CUDA extension rebuilt for sm_100a
                           (oracle/_ref).  This is the stock code path: nothing of ours is on it.
  load_stack("dropin")     stack B: the reference's own models/*.py after
                           spacap3d_b200.install_as_reference_modules(), i.e. the reference's callers
                           (models/backbone_module.py:28-66,101-123, models/voting_module.py,
                           models/proposal_module.py:34-41,57-158, models/SpaCapNet.py:47-74)
                           running unmodified on this package's modules and kernels.

Both bindings want the same module names (`lib.pointnet2.pointnet2_modules`, `pointnet2_utils`,
`models.backbone_module`, ...), so every load happens inside a scrubbed `sys.modules` window and
the imported module objects are detached afterwards; a stack is a plain namespace holding the
classes, which keep working because functions hold their module globals.
"""
import importlib
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
DST = os.path.join(ROOT, "baseline", "_ref")

_TOP = ("lib", "models", "utils", "data", "pointnet2", "pointnet2_utils", "pytorch_utils", "pointnet2_modules",
        "easydict")
CHECKPOINTS = {1: "PRETRAIN_VOTENET_XYZ", 7: "PRETRAIN_VOTENET_XYZ_COLOR_NORMAL",
               132: "PRETRAIN_VOTENET_XYZ_MULTIVIEW_NORMAL"}
_cache = {}


def available():
    from .stage_reference import staged
    return staged()


def _ours(name):
    return name.split(".")[0] in _TOP


class _scrubbed_imports:
    def __enter__(self):
        self.saved = {k: v for k, v in sys.modules.items() if _ours(k)}
        for k in self.saved:
            del sys.modules[k]
        self.path = list(sys.path)
        sys.path[:0] = [DST, os.path.join(DST, "_shims")]
        return self

    def __exit__(self, *exc):
        self.loaded = {k: v for k, v in sys.modules.items() if _ours(k)}
        for k in self.loaded:
            del sys.modules[k]
        sys.modules.update(self.saved)
        sys.path[:] = self.path


def load_stack(binding):
    """-> namespace(SpaCapNet, Pointnet2Backbone, VotingModule, ProposalModule, DC, get_3d_box_batch,
    pointnet2_modules, pointnet2_utils, binding)."""
    if binding in _cache:
        return _cache[binding]
    assert binding in ("reference", "dropin"), binding
    from .stage_reference import stage
    if stage() is None:
        raise FileNotFoundError("baseline/_ref is not staged and /root/reference is absent: run "
                                "`python oracle/stage_reference.py` in the build container")
    with _scrubbed_imports():
        conf = importlib.import_module("lib.config").CONF
        # lib/config.py:9 hard-codes the author's home directory ("TODO: change this"): re-point the two
        # entries the detector reads (data/scannet/model_util_scannet.py:90,101)
        conf.PATH.BASE = DST
        conf.PATH.DATA = os.path.join(DST, "data")
        conf.PATH.SCANNET = os.path.join(DST, "data", "scannet")
        if binding == "reference":
            from .build_ref import load_ref
            pkg = types.ModuleType("pointnet2")
            pkg.__path__ = []
            pkg._ext = load_ref()
            sys.modules["pointnet2"] = pkg
            sys.modules["pointnet2._ext"] = pkg._ext
        else:
            import spacap3d_b200
            spacap3d_b200.install_as_reference_modules()
        net = importlib.import_module("models.SpaCapNet")
        prop = importlib.import_module("models.proposal_module")
        ns = types.SimpleNamespace(
            binding=binding, SpaCapNet=net.SpaCapNet, Pointnet2Backbone=net.Pointnet2Backbone,
            VotingModule=net.VotingModule, ProposalModule=net.ProposalModule, DC=prop.DC,
            get_3d_box_batch=prop.get_3d_box_batch,
            pointnet2_modules=sys.modules["lib.pointnet2.pointnet2_modules"],
            pointnet2_utils=sys.modules["pointnet2_utils"])
    if binding == "reference":
        assert ns.pointnet2_modules.__file__.startswith(DST), ns.pointnet2_modules.__file__
    else:
        assert "spacap3d_b200" in ns.pointnet2_modules.__name__, ns.pointnet2_modules.__name__
    _cache[binding] = ns
    return ns


def checkpoint_path(feature_dim):
    return os.path.join(DST, "pretrained", CHECKPOINTS[feature_dim], "model.pth")


def build_detector(stack, feature_dim, device, pretrained=True):
    """The reference's SpaCapNet detection branch (no_caption=True) in eval mode, optionally with the
    reference's pretrained VoteNet weights (scripts/train.py:170-181 loads them the same way)."""
    import torch
    model = stack.SpaCapNet(num_class=stack.DC.num_class, vocabulary=None, num_heading_bin=stack.DC.num_heading_bin,
                            num_size_cluster=stack.DC.num_size_cluster, mean_size_arr=stack.DC.mean_size_arr,
                            input_feature_dim=feature_dim, num_proposal=256, no_caption=True)
    if pretrained:
        sd = torch.load(checkpoint_path(feature_dim), map_location="cpu")
        missing, unexpected = model.load_state_dict(sd, strict=False)
        assert not missing and not unexpected, (missing, unexpected)
    return model.to(device).eval()
