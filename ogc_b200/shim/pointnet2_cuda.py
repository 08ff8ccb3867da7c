"""`pointnet2_cuda` -- the reference's native module name (pointnet2/pointnet2.py:7 `import pointnet2_cuda as
pointnet2`) bound to libogc_b200.so.  Put THIS directory on PYTHONPATH and the reference's own, byte-for-byte
`pointnet2/pointnet2.py` runs on the sm_100a kernels (option B).

Each function has the name and positional arguments of the pybind wrapper it replaces (pointnet2/src/pointnet2_api.cpp:10-25,
argument orders in src/*.cpp); tensors become raw device pointers, the stream is torch's current stream (the reference's
shims use at::cuda::getCurrentCUDAStream(), src/sampling.cpp:18).  Errors raise instead of exit(-1).
"""
import ctypes
import os

import torch

_LIB_PATH = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "csrc", "libogc_b200.so")
if not os.path.exists(_LIB_PATH):
    raise ImportError(f"pointnet2_cuda shim: {_LIB_PATH} is missing (python -m ogc_b200.build)")
_lib = ctypes.CDLL(_LIB_PATH)
_F = ctypes.c_float


def _p(t):
    if not t.is_cuda or not t.is_contiguous():
        raise ValueError("pointnet2_cuda shim: tensors must be contiguous CUDA tensors")
    return ctypes.c_void_p(t.data_ptr())


def _s():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _chk(rc, name):
    if rc:
        raise RuntimeError(f"libogc_b200: {name} failed with status {rc}")   # the reference would print and exit(-1)
    return 1


def furthest_point_sampling_wrapper(b, n, m, points, temp, idx):
    return _chk(_lib.ogc_furthest_point_sampling(b, n, m, _p(points), _p(temp), _p(idx), _s()), "furthest_point_sampling")


def gather_points_wrapper(b, c, n, npoints, points, idx, out):
    return _chk(_lib.ogc_gather_points(b, c, n, npoints, _p(points), _p(idx), _p(out), _s()), "gather_points")


def gather_points_grad_wrapper(b, c, n, npoints, grad_out, idx, grad_points):
    return _chk(_lib.ogc_gather_points_grad(b, c, n, npoints, _p(grad_out), _p(idx), _p(grad_points), _s()), "gather_points_grad")


def knn_wrapper(b, n, m, k, unknown, known, dist2, idx):
    return _chk(_lib.ogc_knn(b, n, m, k, _p(unknown), _p(known), _p(dist2), _p(idx), _s()), "knn")


def three_nn_wrapper(b, n, m, unknown, known, dist2, idx):
    return _chk(_lib.ogc_three_nn(b, n, m, _p(unknown), _p(known), _p(dist2), _p(idx), _s()), "three_nn")


def three_interpolate_wrapper(b, c, m, n, points, idx, weight, out):
    return _chk(_lib.ogc_three_interpolate(b, c, m, n, _p(points), _p(idx), _p(weight), _p(out), _s()), "three_interpolate")


def three_interpolate_grad_wrapper(b, c, n, m, grad_out, idx, weight, grad_points):
    return _chk(_lib.ogc_three_interpolate_grad(b, c, n, m, _p(grad_out), _p(idx), _p(weight), _p(grad_points), _s()),
                "three_interpolate_grad")


def group_points_wrapper(b, c, n, npoints, nsample, points, idx, out):
    return _chk(_lib.ogc_group_points(b, c, n, npoints, nsample, _p(points), _p(idx), _p(out), _s()), "group_points")


def group_points_grad_wrapper(b, c, n, npoints, nsample, grad_out, idx, grad_points):
    return _chk(_lib.ogc_group_points_grad(b, c, n, npoints, nsample, _p(grad_out), _p(idx), _p(grad_points), _s()),
                "group_points_grad")


def ball_query_wrapper(b, n, m, radius, nsample, new_xyz, xyz, idx):
    return _chk(_lib.ogc_ball_query(b, n, m, _F(radius), nsample, _p(new_xyz), _p(xyz), _p(idx), _s()), "ball_query")
