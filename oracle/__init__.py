"""oracle -- CPU restatement of the reference's PointNet++ kernels (numpy in / numpy out).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package (spacap3d_b200) must never import it.

Parity status: PINNED against tests/golden/ref_ops_*.npz (outputs of the reference's own CUDA
extension run on a B200, see oracle/make_golden.py and tests/test_oracle_golden.py).

The arithmetic lives in pointnet2_oracle.c (each function cites the reference file:line it
follows); this module only marshals numpy arrays through ctypes.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle_pointnet2.so")
_lib = None

_f = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_i = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_c = ctypes.c_int


def build(force=False):
    """gcc-compile pointnet2_oracle.c (same flags as oracle/Makefile)."""
    src = os.path.join(_HERE, "pointnet2_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-mfma", "-ffp-contract=off", "-fopenmp", "-shared",
                               "-fPIC", "-Wall", "-o", _SO, src, "-lm"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        L.orc_opt_n_threads.argtypes = [_c]
        L.orc_opt_n_threads.restype = _c
        L.orc_furthest_point_sampling.argtypes = [_c, _c, _c, _f, _i]
        L.orc_gather_points.argtypes = [_c, _c, _c, _c, _f, _i, _f]
        L.orc_gather_points_grad.argtypes = [_c, _c, _c, _c, _f, _i, _f]
        L.orc_ball_query.argtypes = [_c, _c, _c, ctypes.c_float, _c, _f, _f, _i]
        L.orc_group_points.argtypes = [_c, _c, _c, _c, _c, _f, _i, _f]
        L.orc_group_points_grad.argtypes = [_c, _c, _c, _c, _c, _f, _i, _f]
        L.orc_three_nn.argtypes = [_c, _c, _c, _f, _f, _f, _i]
        L.orc_three_interpolate.argtypes = [_c, _c, _c, _c, _f, _i, _f, _f]
        L.orc_three_interpolate_grad.argtypes = [_c, _c, _c, _c, _f, _i, _f, _f]
        _lib = L
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def opt_n_threads(n):
    return int(lib().orc_opt_n_threads(int(n)))


def furthest_point_sampling(xyz, npoint):
    """xyz (B,N,3) f32 -> (B,npoint) i32.   sampling_gpu.cu:69-173"""
    xyz = _f32(xyz)
    B, N, _ = xyz.shape
    out = np.zeros((B, npoint), np.int32)
    lib().orc_furthest_point_sampling(B, N, npoint, xyz, out)
    return out


def gather_points(points, idx):
    """points (B,C,N), idx (B,M) -> (B,C,M).   sampling_gpu.cu:8-20"""
    points, idx = _f32(points), _i32(idx)
    B, C, N = points.shape
    M = idx.shape[1]
    out = np.zeros((B, C, M), np.float32)
    lib().orc_gather_points(B, C, N, M, points, idx, out)
    return out


def gather_points_grad(grad_out, idx, n):
    """grad_out (B,C,M), idx (B,M) -> (B,C,n).   sampling_gpu.cu:34-47"""
    grad_out, idx = _f32(grad_out), _i32(idx)
    B, C, M = grad_out.shape
    out = np.zeros((B, C, n), np.float32)
    lib().orc_gather_points_grad(B, C, n, M, grad_out, idx, out)
    return out


def ball_query(new_xyz, xyz, radius, nsample):
    """new_xyz (B,M,3), xyz (B,N,3) -> (B,M,nsample) i32.   ball_query_gpu.cu:9-44"""
    new_xyz, xyz = _f32(new_xyz), _f32(xyz)
    B, M, _ = new_xyz.shape
    N = xyz.shape[1]
    out = np.zeros((B, M, nsample), np.int32)
    lib().orc_ball_query(B, N, M, float(radius), int(nsample), new_xyz, xyz, out)
    return out


def group_points(points, idx):
    """points (B,C,N), idx (B,np,ns) -> (B,C,np,ns).   group_points_gpu.cu:8-28"""
    points, idx = _f32(points), _i32(idx)
    B, C, N = points.shape
    _, npoint, ns = idx.shape
    out = np.zeros((B, C, npoint, ns), np.float32)
    lib().orc_group_points(B, C, N, npoint, ns, points, idx, out)
    return out


def group_points_grad(grad_out, idx, n):
    """grad_out (B,C,np,ns), idx (B,np,ns) -> (B,C,n).   group_points_gpu.cu:43-64"""
    grad_out, idx = _f32(grad_out), _i32(idx)
    B, C, npoint, ns = grad_out.shape
    out = np.zeros((B, C, n), np.float32)
    lib().orc_group_points_grad(B, C, n, npoint, ns, grad_out, idx, out)
    return out


def three_nn(unknown, known):
    """unknown (B,n,3), known (B,m,3) -> dist2 (B,n,3) f32, idx (B,n,3) i32.
    interpolate_gpu.cu:9-59 (squared distances, like the C++ entry point)."""
    unknown, known = _f32(unknown), _f32(known)
    B, n, _ = unknown.shape
    m = known.shape[1]
    d2 = np.zeros((B, n, 3), np.float32)
    idx = np.zeros((B, n, 3), np.int32)
    lib().orc_three_nn(B, n, m, unknown, known, d2, idx)
    return d2, idx


def three_interpolate(points, idx, weight):
    """points (B,C,m), idx/weight (B,n,3) -> (B,C,n).   interpolate_gpu.cu:72-101"""
    points, idx, weight = _f32(points), _i32(idx), _f32(weight)
    B, C, m = points.shape
    n = idx.shape[1]
    out = np.zeros((B, C, n), np.float32)
    lib().orc_three_interpolate(B, C, m, n, points, idx, weight, out)
    return out


def three_interpolate_grad(grad_out, idx, weight, m):
    """grad_out (B,C,n), idx/weight (B,n,3) -> (B,C,m).   interpolate_gpu.cu:116-143"""
    grad_out, idx, weight = _f32(grad_out), _i32(idx), _f32(weight)
    B, C, n = grad_out.shape
    out = np.zeros((B, C, m), np.float32)
    lib().orc_three_interpolate_grad(B, C, n, m, grad_out, idx, weight, out)
    return out
