"""Build the UNMODIFIED reference CUDA extension for sm_100a into oracle/_ref/.

TEST INFRASTRUCTURE ONLY.  Nothing in the product path (spacap3d_b200/) may import this.

The reference's nine point-set kernels live in
/root/reference/lib/pointnet2/_ext_src/{src,include} (bindings.cpp:6-19).  Its own setup.py
cannot be used (arch list "3.7+PTX;..." at lib/pointnet2/setup.py:17 is rejected by CUDA 12.x),
so the sources are compiled *where they lie* with torch.utils.cpp_extension.load and
TORCH_CUDA_ARCH_LIST=10.0a.  No reference source is copied into this repository; only the
resulting shared object lands in oracle/_ref/ (git-ignored, but shipped to the GPU box).

Usage:  python oracle/build_ref.py            # no-op if the .so is already there
        python oracle/build_ref.py --force
"""
import glob
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
NAME = "pointnet2_ref_ext"
REF_SRC = "/root/reference/lib/pointnet2/_ext_src"


def ref_so_path():
    return os.path.join(OUT, NAME + ".so")


def build(force=False, verbose=False):
    so = ref_so_path()
    if os.path.exists(so) and not force:
        return so
    if not os.path.isdir(REF_SRC):
        return None  # GPU box: the reference tree is not there; only a prebuilt .so can be used
    os.makedirs(OUT, exist_ok=True)
    os.environ["TORCH_CUDA_ARCH_LIST"] = "10.0a"
    from torch.utils.cpp_extension import load
    sources = sorted(glob.glob(os.path.join(REF_SRC, "src", "*.cpp")) +
                     glob.glob(os.path.join(REF_SRC, "src", "*.cu")))
    load(name=NAME, sources=sources,
         extra_include_paths=[os.path.join(REF_SRC, "include")],
         extra_cflags=["-O2"], extra_cuda_cflags=["-O3"],
         build_directory=OUT, is_python_module=False, verbose=verbose)
    return so if os.path.exists(so) else None


def load_ref():
    """Import the prebuilt reference extension (needs a GPU to *run* anything)."""
    import importlib.util
    import torch  # noqa: F401  (libtorch must be loaded first)
    so = ref_so_path()
    if not os.path.exists(so):
        raise FileNotFoundError(so + " missing: run `python oracle/build_ref.py` in the build container")
    spec = importlib.util.spec_from_file_location(NAME, so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose=True)
    print("reference extension:", p)
