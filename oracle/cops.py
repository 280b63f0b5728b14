"""ctypes bindings of oracle/_build/liboracle.so (oracle/pointnet2_ref.c) on CPU torch tensors.

TEST INFRASTRUCTURE.  Function names and argument order follow the `pointnet2_ops._ext` calls of
the reference (ptt/models/backbones_3d/pointnet2/pointnet2_utils.py:48-287).
"""
import ctypes
import os
import subprocess

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "pointnet2_ref.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = ctypes.CDLL(_LIB_PATH)
    return _lib


def _fp(t):
    assert t.dtype == torch.float32 and t.is_contiguous() and t.device.type == "cpu"
    return ctypes.cast(t.data_ptr(), ctypes.POINTER(ctypes.c_float))


def _ip(t):
    assert t.dtype == torch.int32 and t.is_contiguous() and t.device.type == "cpu"
    return ctypes.cast(t.data_ptr(), ctypes.POINTER(ctypes.c_int))


def furthest_point_sampling(xyz, npoint):
    B, N, _ = xyz.shape
    out = torch.zeros(B, npoint, dtype=torch.int32)
    lib().ref_furthest_point_sampling(B, N, int(npoint), _fp(xyz), _ip(out))
    return out


def furthest_point_sampling_with_dist(dist, npoint):
    B, N, _ = dist.shape
    out = torch.zeros(B, npoint, dtype=torch.int32)
    lib().ref_furthest_point_sampling_with_dist(B, N, int(npoint), _fp(dist), _ip(out))
    return out


def gather_points(points, idx):
    B, C, N = points.shape
    M = idx.shape[1]
    out = torch.empty(B, C, M, dtype=torch.float32)
    lib().ref_gather_points(B, C, N, M, _fp(points), _ip(idx), _fp(out))
    return out


def gather_points_grad(grad_out, idx, n):
    B, C, M = grad_out.shape
    out = torch.empty(B, C, n, dtype=torch.float32)
    lib().ref_gather_points_grad(B, C, int(n), M, _fp(grad_out), _ip(idx), _fp(out))
    return out


def ball_query(new_xyz, xyz, radius, nsample):
    B, N, _ = xyz.shape
    M = new_xyz.shape[1]
    out = torch.empty(B, M, nsample, dtype=torch.int32)
    lib().ref_ball_query(B, N, M, ctypes.c_float(radius), int(nsample), _fp(new_xyz), _fp(xyz), _ip(out))
    return out


def group_points(points, idx):
    B, C, N = points.shape
    _, M, K = idx.shape
    out = torch.empty(B, C, M, K, dtype=torch.float32)
    lib().ref_group_points(B, C, N, M, K, _fp(points), _ip(idx), _fp(out))
    return out


def group_points_grad(grad_out, idx, n):
    B, C, M, K = grad_out.shape
    out = torch.empty(B, C, n, dtype=torch.float32)
    lib().ref_group_points_grad(B, C, int(n), M, K, _fp(grad_out), _ip(idx), _fp(out))
    return out


def three_nn(unknown, known):
    B, n, _ = unknown.shape
    m = known.shape[1]
    dist2 = torch.empty(B, n, 3, dtype=torch.float32)
    idx = torch.empty(B, n, 3, dtype=torch.int32)
    lib().ref_three_nn(B, n, m, _fp(unknown), _fp(known), _fp(dist2), _ip(idx))
    return dist2, idx


def three_interpolate(points, idx, weight):
    B, c, m = points.shape
    n = idx.shape[1]
    out = torch.empty(B, c, n, dtype=torch.float32)
    lib().ref_three_interpolate(B, c, m, n, _fp(points), _ip(idx), _fp(weight), _fp(out))
    return out


def three_interpolate_grad(grad_out, idx, weight, m):
    B, c, n = grad_out.shape
    out = torch.empty(B, c, m, dtype=torch.float32)
    lib().ref_three_interpolate_grad(B, c, n, int(m), _fp(grad_out), _ip(idx), _fp(weight), _fp(out))
    return out


def knn(xyz, k):
    B, n, _ = xyz.shape
    out = torch.empty(B, n, k, dtype=torch.int32)
    lib().ref_knn(B, n, int(k), _fp(xyz), _ip(out))
    return out
