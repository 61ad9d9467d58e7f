"""spacap3d_b200 -- B200-native (sm_100a) PointNet++/VoteNet point-set operators behind the
reference's `pointnet2_utils` / `pointnet2_modules` / `pytorch_utils` API (SpaCap3D's detector
hot path).  See DESIGN.md for the path, its boundary and the kernels; INTEGRATION.md for how the
reference binds to it.
"""
from . import _lib  # noqa: F401


def install_as_reference_modules():
    """Register this package's modules under the names the reference imports
    (`pointnet2._ext`, bare `pointnet2_utils` / `pytorch_utils` / `pointnet2_modules`, and
    `lib.pointnet2.*`), so models/backbone_module.py etc. run unmodified."""
    import sys
    import types
    from . import _ext, pointnet2_modules, pointnet2_utils, pytorch_utils
    pkg = types.ModuleType("pointnet2")
    pkg._ext = _ext
    pkg.__path__ = []
    sys.modules.setdefault("pointnet2", pkg)
    sys.modules.setdefault("pointnet2._ext", _ext)
    for name, mod in (("pointnet2_utils", pointnet2_utils), ("pytorch_utils", pytorch_utils),
                      ("pointnet2_modules", pointnet2_modules)):
        sys.modules.setdefault(name, mod)
        sys.modules.setdefault("lib.pointnet2." + name, mod)
    if "lib" not in sys.modules:
        lib = types.ModuleType("lib")
        lib.__path__ = []
        sys.modules["lib"] = lib
    if "lib.pointnet2" not in sys.modules:
        lp = types.ModuleType("lib.pointnet2")
        lp.__path__ = []
        lp.pointnet2_utils, lp.pytorch_utils, lp.pointnet2_modules = \
            pointnet2_utils, pytorch_utils, pointnet2_modules
        sys.modules["lib.pointnet2"] = lp
        sys.modules["lib"].pointnet2 = lp
