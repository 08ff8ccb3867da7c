"""TEST INFRASTRUCTURE -- the REFERENCE's own CUDA extension (built by oracle/ref_build.py into
oracle/_ref/pointnet2_cuda_ref.so) behind the same back-end interface as the product and the CPU
oracle.  Used on the GPU box to (a) validate the CPU restatement against the real thing,
(b) check the sm_100a kernels against the real thing, (c) time the reference extension in bench.py.

Allocation / pre-fill conventions are the reference's (pointnet2/pointnet2.py:32-33,61,99-100,
130-131,163,206,251).
"""
import importlib.util
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "_ref", "pointnet2_cuda_ref.so")
_mod = None


def available() -> bool:
    return os.path.exists(SO_PATH)


def module():
    global _mod
    if _mod is None:
        spec = importlib.util.spec_from_file_location("pointnet2_cuda_ref", SO_PATH)
        _mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(_mod)
    return _mod


class RefExtBackend:
    name = "refext"
    launches = 0   # interface parity with B200Backend (bench.py counts OUR kernels only)

    def __init__(self):
        self.m = module()

    def fps(self, xyz, npoint):
        B, N, _ = xyz.shape
        out = torch.empty(B, npoint, dtype=torch.int32, device=xyz.device)
        temp = torch.full((B, N), 1e10, dtype=torch.float32, device=xyz.device)
        self.m.furthest_point_sampling_wrapper(B, N, npoint, xyz, temp, out)
        return out

    def knn(self, k, unknown, known):
        B, n, _ = unknown.shape
        m = known.shape[1]
        d2 = torch.empty(B, n, k, dtype=torch.float32, device=unknown.device)
        idx = torch.empty(B, n, k, dtype=torch.int32, device=unknown.device)
        self.m.knn_wrapper(B, n, m, k, unknown, known, d2, idx)
        return d2, idx

    def three_nn(self, unknown, known):
        B, n, _ = unknown.shape
        m = known.shape[1]
        d2 = torch.empty(B, n, 3, dtype=torch.float32, device=unknown.device)
        idx = torch.empty(B, n, 3, dtype=torch.int32, device=unknown.device)
        self.m.three_nn_wrapper(B, n, m, unknown, known, d2, idx)
        return d2, idx

    def three_interpolate(self, features, idx, weight):
        B, c, m = features.shape
        n = idx.shape[1]
        out = torch.empty(B, c, n, dtype=torch.float32, device=features.device)
        self.m.three_interpolate_wrapper(B, c, m, n, features, idx, weight, out)
        return out

    def three_interpolate_grad(self, grad_out, idx, weight, m):
        B, c, n = grad_out.shape
        g = torch.zeros(B, c, m, dtype=torch.float32, device=grad_out.device)
        self.m.three_interpolate_grad_wrapper(B, c, n, m, grad_out, idx, weight, g)
        return g

    def group_points(self, features, idx):
        B, C, N = features.shape
        _, M, S = idx.shape
        out = torch.empty(B, C, M, S, dtype=torch.float32, device=features.device)
        self.m.group_points_wrapper(B, C, N, M, S, features, idx, out)
        return out

    def group_points_grad(self, grad_out, idx, N):
        B, C, M, S = grad_out.shape
        g = torch.zeros(B, C, N, dtype=torch.float32, device=grad_out.device)
        self.m.group_points_grad_wrapper(B, C, N, M, S, grad_out, idx, g)
        return g

    def gather_points(self, features, idx):
        B, C, N = features.shape
        M = idx.shape[1]
        out = torch.empty(B, C, M, dtype=torch.float32, device=features.device)
        self.m.gather_points_wrapper(B, C, N, M, features, idx, out)
        return out

    def gather_points_grad(self, grad_out, idx, N):
        B, C, M = grad_out.shape
        g = torch.zeros(B, C, N, dtype=torch.float32, device=grad_out.device)
        self.m.gather_points_grad_wrapper(B, C, N, M, grad_out, idx, g)
        return g

    def ball_query(self, radius, nsample, xyz, new_xyz):
        B, N, _ = xyz.shape
        M = new_xyz.shape[1]
        idx = torch.zeros(B, M, nsample, dtype=torch.int32, device=xyz.device)
        self.m.ball_query_wrapper(B, N, M, float(radius), nsample, new_xyz, xyz, idx)
        return idx
