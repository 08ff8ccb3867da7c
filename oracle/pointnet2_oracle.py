"""TEST INFRASTRUCTURE -- ctypes front-end of the CPU oracle (oracle/pointnet2_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module.  It exposes the ten native entry points of the reference
`pointnet2_cuda` module (pointnet2/src/pointnet2_api.cpp:10-25) with the same names,
argument order and caller-allocated outputs, operating on CPU torch tensors, plus
`OracleBackend`, the object tests install behind `pointnet2.pointnet2` to run the
operator API without a GPU.

Parity status: pinned differentially only (the reference has no tests / golden vectors):
see the header of pointnet2_oracle.c.
"""
import ctypes
import os
import sys

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
if _HERE not in sys.path:
    sys.path.insert(0, _HERE)

_lib = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "_build", "liboracle.so")
        if not os.path.exists(path):
            import importlib.util
            spec = importlib.util.spec_from_file_location("oracle_build", os.path.join(_HERE, "build.py"))
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            path = mod.build()
        _lib = ctypes.CDLL(path)
        _lib.oracle_opt_n_threads.restype = ctypes.c_int
        _lib.oracle_knn.restype = ctypes.c_int
    return _lib


def _f(t):
    assert t.dtype == torch.float32 and t.is_contiguous() and t.device.type == "cpu"
    return ctypes.c_void_p(t.data_ptr())


def _i(t):
    assert t.dtype == torch.int32 and t.is_contiguous() and t.device.type == "cpu"
    return ctypes.c_void_p(t.data_ptr())


def opt_n_threads(n: int) -> int:
    return lib().oracle_opt_n_threads(int(n))


# ---- the ten wrappers, reference names/argument order (pointnet2/src/*.cpp) -----------------

def furthest_point_sampling_wrapper(b, n, m, points, temp, idx):
    lib().oracle_furthest_point_sampling(b, n, m, _f(points), _f(temp), _i(idx))
    return 1


def gather_points_wrapper(b, c, n, npoints, points, idx, out):
    lib().oracle_gather_points(b, c, n, npoints, _f(points), _i(idx), _f(out))
    return 1


def gather_points_grad_wrapper(b, c, n, npoints, grad_out, idx, grad_points):
    lib().oracle_gather_points_grad(b, c, n, npoints, _f(grad_out), _i(idx), _f(grad_points))
    return 1


def knn_wrapper(b, n, m, k, unknown, known, dist2, idx):
    rc = lib().oracle_knn(b, n, m, k, _f(unknown), _f(known), _f(dist2), _i(idx))
    if rc != 0:
        raise ValueError("oracle_knn: k must be in [0, 200]")
    return 1


def three_nn_wrapper(b, n, m, unknown, known, dist2, idx):
    lib().oracle_three_nn(b, n, m, _f(unknown), _f(known), _f(dist2), _i(idx))
    return 1


def three_interpolate_wrapper(b, c, m, n, points, idx, weight, out):
    lib().oracle_three_interpolate(b, c, m, n, _f(points), _i(idx), _f(weight), _f(out))
    return 1


def three_interpolate_grad_wrapper(b, c, n, m, grad_out, idx, weight, grad_points):
    lib().oracle_three_interpolate_grad(b, c, n, m, _f(grad_out), _i(idx), _f(weight), _f(grad_points))
    return 1


def group_points_wrapper(b, c, n, npoints, nsample, points, idx, out):
    lib().oracle_group_points(b, c, n, npoints, nsample, _f(points), _i(idx), _f(out))
    return 1


def group_points_grad_wrapper(b, c, n, npoints, nsample, grad_out, idx, grad_points):
    lib().oracle_group_points_grad(b, c, n, npoints, nsample, _f(grad_out), _i(idx), _f(grad_points))
    return 1


def ball_query_wrapper(b, n, m, radius, nsample, new_xyz, xyz, idx):
    lib().oracle_ball_query(b, n, m, ctypes.c_float(radius), nsample, _f(new_xyz), _f(xyz), _i(idx))
    return 1


class OracleBackend:
    """Same method set as ogc_b200.backend.B200Backend, on CPU tensors.

    Allocation / pre-fill conventions follow pointnet2/pointnet2.py of the reference:
    temp = 1e10 (:33), ball-query idx pre-zeroed (:251), grads zero-initialised
    (:72,181,223).  Distances are returned squared, as the native ops do; the sqrt lives in
    the operator layer (:103,134).
    """
    name = "oracle"
    launches = 0   # interface parity with B200Backend (bench.py counts OUR kernels only)

    def fps(self, xyz, npoint):
        B, N, _ = xyz.shape
        out = torch.empty(B, npoint, dtype=torch.int32)
        temp = torch.full((B, N), 1e10, dtype=torch.float32)
        furthest_point_sampling_wrapper(B, N, npoint, xyz, temp, out)
        return out

    def knn(self, k, unknown, known):
        B, n, _ = unknown.shape
        m = known.shape[1]
        d2 = torch.empty(B, n, k, dtype=torch.float32)
        idx = torch.empty(B, n, k, dtype=torch.int32)
        knn_wrapper(B, n, m, k, unknown, known, d2, idx)
        return d2, idx

    def three_nn(self, unknown, known):
        B, n, _ = unknown.shape
        m = known.shape[1]
        d2 = torch.empty(B, n, 3, dtype=torch.float32)
        idx = torch.empty(B, n, 3, dtype=torch.int32)
        three_nn_wrapper(B, n, m, unknown, known, d2, idx)
        return d2, idx

    def three_interpolate(self, features, idx, weight):
        B, c, m = features.shape
        n = idx.shape[1]
        out = torch.empty(B, c, n, dtype=torch.float32)
        three_interpolate_wrapper(B, c, m, n, features, idx, weight, out)
        return out

    def three_interpolate_grad(self, grad_out, idx, weight, m):
        B, c, n = grad_out.shape
        g = torch.zeros(B, c, m, dtype=torch.float32)
        three_interpolate_grad_wrapper(B, c, n, m, grad_out, idx, weight, g)
        return g

    def group_points(self, features, idx):
        B, C, N = features.shape
        _, M, S = idx.shape
        out = torch.empty(B, C, M, S, dtype=torch.float32)
        group_points_wrapper(B, C, N, M, S, features, idx, out)
        return out

    def group_points_grad(self, grad_out, idx, N):
        B, C, M, S = grad_out.shape
        g = torch.zeros(B, C, N, dtype=torch.float32)
        group_points_grad_wrapper(B, C, N, M, S, grad_out, idx, g)
        return g

    def gather_points(self, features, idx):
        B, C, N = features.shape
        M = idx.shape[1]
        out = torch.empty(B, C, M, dtype=torch.float32)
        gather_points_wrapper(B, C, N, M, features, idx, out)
        return out

    def gather_points_grad(self, grad_out, idx, N):
        B, C, M = grad_out.shape
        g = torch.zeros(B, C, N, dtype=torch.float32)
        gather_points_grad_wrapper(B, C, N, M, grad_out, idx, g)
        return g

    def ball_query(self, radius, nsample, xyz, new_xyz):
        B, N, _ = xyz.shape
        M = new_xyz.shape[1]
        idx = torch.zeros(B, M, nsample, dtype=torch.int32)
        ball_query_wrapper(B, N, M, float(radius), nsample, new_xyz, xyz, idx)
        return idx
