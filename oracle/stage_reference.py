"""Stage the reference's UNMODIFIED Python detector stack + three pretrained detector checkpoints
under baseline/_ref/ (git-ignored, shipped to the GPU box by gpurun).

TEST INFRASTRUCTURE ONLY (used by tests/ and bench.py --impl reference; the product package never
touches it).  Nothing staged here enters the repository history: the files are copied byte for
byte from /root/reference at build time, in the build container only (the GPU box has no
/root/reference and uses the staged copy).

What is staged, and why each file is needed to run `models/SpaCapNet.py` (detection branch,
`no_caption=True`) through the reference's own stock code path:

  lib/pointnet2/{pointnet2_utils,pointnet2_modules,pytorch_utils}.py   the SA / FP modules
  models/{SpaCapNet,backbone_module,voting_module,proposal_module,transformer_captioner}.py
  utils/{box_util,nn_distance}.py            get_3d_box_batch (host box decode); captioner import
  data/scannet/model_util_scannet.py         ScannetDatasetConfig (param2obb_batch, mean sizes)
  data/scannet/meta_data/{scannet_reference_means.npz, scannetv2-labels.combined.tsv}
  lib/config.py                              unmodified; CONF.PATH.* is re-pointed at run time
  pretrained/PRETRAIN_VOTENET_XYZ{,_COLOR_NORMAL,_MULTIVIEW_NORMAL}/model.pth   (C = 1 / 7 / 132)

Two shims are WRITTEN (not copied) because the packages are absent from this image:
  _shims/easydict.py   a 10-line attribute dict (lib/config.py:3 imports easydict)

`oracle/refstack.py` imports the staged tree in three bindings (reference ext / our ext / ours).
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
DST = os.path.join(ROOT, "baseline", "_ref")

FILES = [
    "lib/pointnet2/pointnet2_utils.py", "lib/pointnet2/pointnet2_modules.py", "lib/pointnet2/pytorch_utils.py",
    "lib/config.py",
    "models/SpaCapNet.py", "models/backbone_module.py", "models/voting_module.py", "models/proposal_module.py",
    "models/transformer_captioner.py",
    "utils/box_util.py", "utils/nn_distance.py",
    "data/scannet/model_util_scannet.py",
    "data/scannet/meta_data/scannet_reference_means.npz",
    "data/scannet/meta_data/scannetv2-labels.combined.tsv",
    "pretrained/PRETRAIN_VOTENET_XYZ/model.pth",
    "pretrained/PRETRAIN_VOTENET_XYZ_COLOR_NORMAL/model.pth",
    "pretrained/PRETRAIN_VOTENET_XYZ_MULTIVIEW_NORMAL/model.pth",
]

EASYDICT_SHIM = '''"""Minimal stand-in for the `easydict` package (absent from this image): attribute access on a dict,
nested dicts converted on assignment.  Written by oracle/stage_reference.py; not reference code."""


class EasyDict(dict):
    def __init__(self, d=None, **kw):
        super().__init__()
        for k, v in dict(d or {}, **kw).items():
            setattr(self, k, v)

    def __setattr__(self, k, v):
        if isinstance(v, dict) and not isinstance(v, EasyDict):
            v = EasyDict(v)
        super().__setitem__(k, v)
        super().__setattr__(k, v)

    __setitem__ = __setattr__
'''


def staged():
    return all(os.path.exists(os.path.join(DST, f)) for f in FILES) and \
        os.path.exists(os.path.join(DST, "_shims", "easydict.py"))


def stage(force=False):
    """Copy the files (build container only).  Returns baseline/_ref or None when neither the staged
    tree nor /root/reference is available."""
    if staged() and not force:
        return DST
    if not os.path.isdir(REF):
        return DST if staged() else None
    for f in FILES:
        dst = os.path.join(DST, f)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(REF, f), dst)
    os.makedirs(os.path.join(DST, "_shims"), exist_ok=True)
    with open(os.path.join(DST, "_shims", "easydict.py"), "w") as fh:
        fh.write(EASYDICT_SHIM)
    return DST


if __name__ == "__main__":
    print("staged reference stack:", stage(force="--force" in sys.argv))
