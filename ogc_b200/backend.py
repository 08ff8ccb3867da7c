"""Operator back-ends behind `pointnet2.pointnet2`.

`B200Backend` is the product: it allocates outputs with torch (caching allocator), passes raw
device pointers + the CURRENT CUDA stream to the C ABI of libogc_b200.so (include/ogc_b200.h) and
never synchronises.  Method set = the ten native entry points of the reference's
`pointnet2_cuda` module (pointnet2/src/pointnet2_api.cpp:10-25), with allocation folded in.

`set_backend()` is the seam tests use to run the same operator layer on the CPU oracle or on the
reference extension; product code never imports `oracle/`.
"""
import ctypes
from contextlib import contextmanager

import torch

from . import _lib


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _chk_f32(t, name):
    if not (isinstance(t, torch.Tensor) and t.is_cuda):
        raise RuntimeError(f"ogc_b200: `{name}` must be a CUDA tensor (no CPU fallback exists); got "
                           f"{t.device if isinstance(t, torch.Tensor) else type(t)}")
    if t.dtype != torch.float32:
        raise ValueError(f"ogc_b200: `{name}` must be float32, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"ogc_b200: `{name}` must be contiguous")


def _chk_i32(t, name):
    if not (isinstance(t, torch.Tensor) and t.is_cuda):
        raise RuntimeError(f"ogc_b200: `{name}` must be a CUDA tensor (no CPU fallback exists)")
    if t.dtype != torch.int32:
        raise ValueError(f"ogc_b200: `{name}` must be int32, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"ogc_b200: `{name}` must be contiguous")


class OpTimer:
    """Optional per-op CUDA-event timing on the launching stream (used by bench.py for the
    roofline block).  Disabled by default: zero overhead beyond one attribute test."""

    def __init__(self):
        self.enabled = False
        self.detail = False   # per-shape span names (profiling scripts)
        self.records = []  # (name, start_event, end_event, algorithmic_bytes, useful_flops)

    @contextmanager
    def span(self, name, nbytes, flops=0):
        if not self.enabled:
            yield
            return
        s = torch.cuda.Event(enable_timing=True)
        e = torch.cuda.Event(enable_timing=True)
        s.record()
        yield
        e.record()
        self.records.append((name, s, e, nbytes, flops))

    def summary(self):
        """{name: {"calls", "ms", "bytes", "flops"}} -- call after torch.cuda.synchronize().  `flops` = useful
        fp32-equivalent FLOPs of a contraction kernel (0 for data-movement kernels)."""
        out = {}
        for name, s, e, nb, fl in self.records:
            d = out.setdefault(name, {"calls": 0, "ms": 0.0, "bytes": 0, "flops": 0})
            d["calls"] += 1
            d["ms"] += s.elapsed_time(e)
            d["bytes"] += nb
            d["flops"] += fl
        return out

    def reset(self):
        self.records = []


TIMER = OpTimer()


class B200Backend:
    name = "b200"

    def __init__(self):
        self.lib = _lib.load()  # raises if the extension is missing
        self.launches = 0       # number of libogc_b200 kernels launched (bench.py's gpu_launches)

    # ---- K1 ------------------------------------------------------------------------------
    def fps(self, xyz, npoint):
        _chk_f32(xyz, "xyz")
        B, N, _ = xyz.shape
        out = torch.empty(B, npoint, dtype=torch.int32, device=xyz.device)
        temp = torch.empty(B, N, dtype=torch.float32, device=xyz.device) if N > 16384 else None
        with TIMER.span("fps", B * (12 * N + 4 * npoint)):
            _lib.check(self.lib.ogc_furthest_point_sampling(B, N, npoint, _ptr(xyz), _ptr(temp), _ptr(out),
                                                            _stream()), "ogc_furthest_point_sampling")
        self.launches += 1
        return out

    # ---- K4 / K5 ---------------------------------------------------------------------------
    def knn(self, k, unknown, known, sqrt=False):
        _chk_f32(unknown, "unknown")
        _chk_f32(known, "known")
        B, n, _ = unknown.shape
        m = known.shape[1]
        d = torch.empty(B, n, k, dtype=torch.float32, device=unknown.device)
        idx = torch.empty(B, n, k, dtype=torch.int32, device=unknown.device)
        fn = self.lib.ogc_knn_sqrt if sqrt else self.lib.ogc_knn
        with TIMER.span("knn", B * (12 * (n + m) + 8 * n * k)):
            _lib.check(fn(B, n, m, k, _ptr(unknown), _ptr(known), _ptr(d), _ptr(idx), _stream()), "ogc_knn")
        self.launches += 1
        return d, idx

    # ---- uniform-grid acceleration of the radius-bounded searches (csrc/grid.cu) ----------------
    USE_GRID = True          # tests flip this to compare with the brute-force kernels
    GRID_MIN_POINTS = 1024   # below this the brute-force kernels are already cheap

    def _grid(self, xyz, radius):
        """Counting-sort every cloud of xyz (B,m,3) into a grid with cells >= radius -> (sorted, cell_start, params)."""
        B, m, _ = xyz.shape
        dev = xyz.device
        sorted_ = torch.empty(B, m, 4, dtype=torch.float32, device=dev)
        table = torch.empty(self.lib.ogc_grid_table_bytes(B) // 4, dtype=torch.int32, device=dev)
        params = torch.empty(B, 16, dtype=torch.float32, device=dev)
        with TIMER.span("grid_build", B * m * 28):
            _lib.check(self.lib.ogc_grid_build(B, m, float(radius), _ptr(xyz), _ptr(sorted_), _ptr(table), _ptr(params),
                                               _stream()), "ogc_grid_build")
        self.launches += 1
        return sorted_, table, params

    @staticmethod
    def _same(a, b):
        return a.data_ptr() == b.data_ptr() and a.shape == b.shape

    def knn_bounded(self, k, unknown, known, max_dist):
        """k-NN among candidates within max_dist (dist = sqrt, +inf / idx 0 for unfilled slots); see the header."""
        _chk_f32(unknown, "unknown")
        _chk_f32(known, "known")
        B, n, _ = unknown.shape
        m = known.shape[1]
        d = torch.empty(B, n, k, dtype=torch.float32, device=unknown.device)
        idx = torch.empty(B, n, k, dtype=torch.int32, device=unknown.device)
        if self.USE_GRID and m >= self.GRID_MIN_POINTS and max_dist > 0:
            sorted_, table, params = self._grid(known, max_dist)
            with TIMER.span("knn", B * (12 * (n + m) + 8 * n * k)):
                _lib.check(self.lib.ogc_knn_grid(B, n, m, k, float(max_dist), None if self._same(unknown, known) else _ptr(unknown),
                                                 _ptr(sorted_), _ptr(table), _ptr(params), _ptr(d), _ptr(idx), _stream()),
                           "ogc_knn_grid")
            self.launches += 1
            return d, idx
        with TIMER.span("knn", B * (12 * (n + m) + 8 * n * k)):
            _lib.check(self.lib.ogc_knn_bounded(B, n, m, k, float(max_dist), _ptr(unknown), _ptr(known), _ptr(d),
                                                _ptr(idx), _stream()), "ogc_knn_bounded")
        self.launches += 1
        return d, idx

    def three_nn(self, unknown, known):
        _chk_f32(unknown, "unknown")
        _chk_f32(known, "known")
        B, n, _ = unknown.shape
        m = known.shape[1]
        d2 = torch.empty(B, n, 3, dtype=torch.float32, device=unknown.device)
        idx = torch.empty(B, n, 3, dtype=torch.int32, device=unknown.device)
        with TIMER.span("three_nn", B * (12 * (n + m) + 24 * n)):
            _lib.check(self.lib.ogc_three_nn(B, n, m, _ptr(unknown), _ptr(known), _ptr(d2), _ptr(idx), _stream()),
                       "ogc_three_nn")
        self.launches += 1
        return d2, idx

    # ---- K6 / K7 ---------------------------------------------------------------------------
    def three_interpolate(self, features, idx, weight):
        _chk_f32(features, "features")
        _chk_i32(idx, "idx")
        _chk_f32(weight, "weight")
        B, c, m = features.shape
        n = idx.shape[1]
        out = torch.empty(B, c, n, dtype=torch.float32, device=features.device)
        with TIMER.span("three_interpolate", B * (4 * c * m + 24 * n + 4 * c * n)):
            _lib.check(self.lib.ogc_three_interpolate(B, c, m, n, _ptr(features), _ptr(idx), _ptr(weight),
                                                      _ptr(out), _stream()), "ogc_three_interpolate")
        self.launches += 1
        return out

    def three_interpolate_grad(self, grad_out, idx, weight, m):
        _chk_f32(grad_out, "grad_out")
        _chk_i32(idx, "idx")
        _chk_f32(weight, "weight")
        B, c, n = grad_out.shape
        g = torch.zeros(B, c, m, dtype=torch.float32, device=grad_out.device)
        with TIMER.span("three_interpolate_grad", B * (4 * c * m + 24 * n + 4 * c * n)):
            _lib.check(self.lib.ogc_three_interpolate_grad(B, c, n, m, _ptr(grad_out), _ptr(idx), _ptr(weight),
                                                           _ptr(g), _stream()), "ogc_three_interpolate_grad")
        self.launches += 1
        return g

    # ---- K8 / K9 ---------------------------------------------------------------------------
    def group_points(self, features, idx):
        _chk_f32(features, "features")
        _chk_i32(idx, "idx")
        B, C, N = features.shape
        _, M, S = idx.shape
        out = torch.empty(B, C, M, S, dtype=torch.float32, device=features.device)
        with TIMER.span("group_points", B * (4 * C * N + 4 * M * S + 4 * C * M * S)):
            _lib.check(self.lib.ogc_group_points(B, C, N, M, S, _ptr(features), _ptr(idx), _ptr(out), _stream()),
                       "ogc_group_points")
        self.launches += 1
        return out

    def group_points_grad(self, grad_out, idx, N):
        _chk_f32(grad_out, "grad_out")
        _chk_i32(idx, "idx")
        B, C, M, S = grad_out.shape
        g = torch.zeros(B, C, N, dtype=torch.float32, device=grad_out.device)
        with TIMER.span("group_points_grad", B * (4 * C * N + 4 * M * S + 4 * C * M * S)):
            _lib.check(self.lib.ogc_group_points_grad(B, C, N, M, S, _ptr(grad_out), _ptr(idx), _ptr(g),
                                                      _stream()), "ogc_group_points_grad")
        self.launches += 1
        return g

    # ---- K2 / K3 ---------------------------------------------------------------------------
    def gather_points(self, features, idx):
        _chk_f32(features, "features")
        _chk_i32(idx, "idx")
        B, C, N = features.shape
        M = idx.shape[1]
        out = torch.empty(B, C, M, dtype=torch.float32, device=features.device)
        with TIMER.span("gather_points", B * (4 * C * N + 4 * M + 4 * C * M)):
            _lib.check(self.lib.ogc_gather_points(B, C, N, M, _ptr(features), _ptr(idx), _ptr(out), _stream()),
                       "ogc_gather_points")
        self.launches += 1
        return out

    def gather_points_grad(self, grad_out, idx, N):
        _chk_f32(grad_out, "grad_out")
        _chk_i32(idx, "idx")
        B, C, M = grad_out.shape
        g = torch.zeros(B, C, N, dtype=torch.float32, device=grad_out.device)
        with TIMER.span("gather_points_grad", B * (4 * C * N + 4 * M + 4 * C * M)):
            _lib.check(self.lib.ogc_gather_points_grad(B, C, N, M, _ptr(grad_out), _ptr(idx), _ptr(g), _stream()),
                       "ogc_gather_points_grad")
        self.launches += 1
        return g

    # ---- K10 -------------------------------------------------------------------------------
    def ball_query(self, radius, nsample, xyz, new_xyz):
        _chk_f32(xyz, "xyz")
        _chk_f32(new_xyz, "new_xyz")
        B, N, _ = xyz.shape
        M = new_xyz.shape[1]
        idx = torch.empty(B, M, nsample, dtype=torch.int32, device=xyz.device)
        if self.USE_GRID and self.GRID_MIN_POINTS <= N <= 32768 and radius > 0:
            sorted_, table, params = self._grid(xyz, radius)
            with TIMER.span("ball_query", B * (12 * (N + M) + 4 * M * nsample)):
                _lib.check(self.lib.ogc_ball_query_grid(B, M, N, float(radius), nsample,
                                                        None if self._same(new_xyz, xyz) else _ptr(new_xyz), _ptr(sorted_),
                                                        _ptr(table), _ptr(params), _ptr(idx), _stream()), "ogc_ball_query_grid")
            self.launches += 1
            return idx
        with TIMER.span("ball_query", B * (12 * (N + M) + 4 * M * nsample)):
            _lib.check(self.lib.ogc_ball_query(B, N, M, float(radius), nsample, _ptr(new_xyz), _ptr(xyz), _ptr(idx),
                                               _stream()), "ogc_ball_query")
        self.launches += 1
        return idx

    # ---- fused OGC-loss kernels (csrc/losses.cu) -------------------------------------------
    def weighted_kabsch(self, pc, second, mask, second_is_flow=True):
        """pc, second (B,N,3), mask (B,N,K) -> Rt (B,K,12) = [R row-major | t]."""
        _chk_f32(pc, "pc"); _chk_f32(second, "second"); _chk_f32(mask, "mask")
        B, N, K = mask.shape
        Rt = torch.empty(B, K, 12, dtype=torch.float32, device=pc.device)
        with TIMER.span("weighted_kabsch", B * N * (24 + 4 * K) * 2):
            _lib.check(self.lib.ogc_weighted_kabsch(B, N, K, int(bool(second_is_flow)), _ptr(pc), _ptr(second),
                                                    _ptr(mask), _ptr(Rt), _stream()), "ogc_weighted_kabsch")
        self.launches += 1
        return Rt

    def dynamic_loss(self, pc, flow, mask, need_grad=True):
        """-> loss_pt (B,N), grad_mask (B,N,K) or None, Rt (B,K,12)."""
        _chk_f32(pc, "pc"); _chk_f32(flow, "flow"); _chk_f32(mask, "mask")
        B, N, K = mask.shape
        loss_pt = torch.empty(B, N, dtype=torch.float32, device=pc.device)
        grad = torch.empty(B, N, K, dtype=torch.float32, device=pc.device) if need_grad else None
        Rt = torch.empty(B, K, 12, dtype=torch.float32, device=pc.device)
        with TIMER.span("dynamic_loss", B * N * ((24 + 4 * K) * 3 + 4 + 4 * K)):
            _lib.check(self.lib.ogc_dynamic_loss(B, N, K, _ptr(pc), _ptr(flow), _ptr(mask), _ptr(loss_pt),
                                                 _ptr(grad), _ptr(Rt), _stream()), "ogc_dynamic_loss")
        self.launches += 1
        return loss_pt, grad, Rt

    def apply_rigid_flow(self, pc, mask, Rt):
        _chk_f32(pc, "pc"); _chk_f32(mask, "mask"); _chk_f32(Rt, "Rt")
        B, N, K = mask.shape
        out = torch.empty(B, N, 3, dtype=torch.float32, device=pc.device)
        with TIMER.span("apply_rigid_flow", B * N * (24 + 4 * K)):
            _lib.check(self.lib.ogc_apply_rigid_flow(B, N, K, _ptr(pc), _ptr(mask), _ptr(Rt), _ptr(out), _stream()),
                       "ogc_apply_rigid_flow")
        self.launches += 1
        return out

    def neighbor_l1(self, mask, idx, dist, radius, coef, grad_accum=None, need_loss=True):
        """loss_pt (B,N); grad_accum (B,N,K) += coef * d(sum loss_pt)/d mask when given."""
        _chk_f32(mask, "mask"); _chk_i32(idx, "idx")
        if dist is not None:
            _chk_f32(dist, "dist")
        B, N, K = mask.shape
        S = idx.shape[2]
        loss_pt = torch.empty(B, N, dtype=torch.float32, device=mask.device) if need_loss else None
        with TIMER.span("neighbor_l1", B * N * (8 * S + 4 * K * (S + 2))):
            _lib.check(self.lib.ogc_neighbor_l1(B, N, K, S, _ptr(mask), _ptr(idx), _ptr(dist),
                                                float(radius if radius is not None else 0.0), float(coef),
                                                _ptr(loss_pt), _ptr(grad_accum), _stream()), "ogc_neighbor_l1")
        self.launches += 1
        return loss_pt

    def mask_contingency(self, mask1, mask2):
        _chk_f32(mask1, "mask1"); _chk_f32(mask2, "mask2")
        B, N, K = mask1.shape
        inter = torch.zeros(B, K, K, dtype=torch.int32, device=mask1.device)
        with TIMER.span("mask_contingency", B * N * 8 * K):
            _lib.check(self.lib.ogc_mask_contingency(B, N, K, _ptr(mask1), _ptr(mask2), _ptr(inter), _stream()),
                       "ogc_mask_contingency")
        self.launches += 1
        return inter

    def invariance_loss(self, mask1, mask2, perm12, perm21, need_grad=True):
        _chk_f32(mask1, "mask1"); _chk_f32(mask2, "mask2"); _chk_i32(perm12, "perm12"); _chk_i32(perm21, "perm21")
        B, N, K = mask1.shape
        loss_pt = torch.empty(B, N, dtype=torch.float32, device=mask1.device)
        g1 = torch.empty_like(mask1) if need_grad else None
        g2 = torch.empty_like(mask2) if need_grad else None
        with TIMER.span("invariance_loss", B * N * (16 * K + 4)):
            _lib.check(self.lib.ogc_invariance_loss(B, N, K, _ptr(mask1), _ptr(mask2), _ptr(perm12), _ptr(perm21),
                                                    _ptr(loss_pt), _ptr(g1), _ptr(g2), _stream()),
                       "ogc_invariance_loss")
        self.launches += 1
        return loss_pt, g1, g2

    def icp_correspond(self, pc1, flow, pc2, mask1, mask2, temperature):
        for t, nme in ((pc1, "pc1"), (flow, "flow"), (pc2, "pc2"), (mask1, "mask1"), (mask2, "mask2")):
            _chk_f32(t, nme)
        B, N1, K = mask1.shape
        N2 = pc2.shape[1]
        out = torch.empty(B, N1, 3, dtype=torch.float32, device=pc1.device)
        with TIMER.span("icp_correspond", B * (N1 * (36 + 4 * K) + N2 * (12 + 4 * K))):
            _lib.check(self.lib.ogc_icp_correspond(B, N1, N2, K, float(temperature), _ptr(pc1), _ptr(flow), _ptr(pc2),
                                                   _ptr(mask1), _ptr(mask2), _ptr(out), _stream()), "ogc_icp_correspond")
        self.launches += 1
        return out

    def softmax_transfer(self, query, key, val, temperature):
        """out[m] = sum_n softmax_n(-|query_m - key_n| / T) val[n]; query (B,N1,3), key (B,N2,3), val (B,N2,K)."""
        for t, nme in ((query, "query"), (key, "key"), (val, "val")):
            _chk_f32(t, nme)
        B, N1, _ = query.shape
        N2, K = val.shape[1], val.shape[2]
        out = torch.empty(B, N1, K, dtype=torch.float32, device=query.device)
        with TIMER.span("softmax_transfer", B * (N1 * (12 + 4 * K) + N2 * (12 + 4 * K))):
            _lib.check(self.lib.ogc_softmax_transfer(B, N1, N2, K, float(temperature), _ptr(query), _ptr(key), _ptr(val),
                                                     _ptr(out), _stream()), "ogc_softmax_transfer")
        self.launches += 1
        return out

    def mask_match(self, inter):
        """inter (B,K,K) int32 -> (perm12, perm21) (B,K) int32, Hungarian on the device (no host sync)."""
        _chk_i32(inter, "inter")
        B, K, _ = inter.shape
        p12 = torch.empty(B, K, dtype=torch.int32, device=inter.device)
        p21 = torch.empty(B, K, dtype=torch.int32, device=inter.device)
        with TIMER.span("mask_match", B * K * K * 4):
            _lib.check(self.lib.ogc_mask_match(B, K, _ptr(inter), _ptr(p12), _ptr(p21), _stream()), "ogc_mask_match")
        self.launches += 1
        return p12, p21

    def mask_nuclear_norm(self, mask):
        _chk_f32(mask, "mask")
        B, N, K = mask.shape
        out = torch.empty(B, dtype=torch.float32, device=mask.device)
        ws = torch.empty(B * 1024, dtype=torch.float64, device=mask.device)
        with TIMER.span("mask_nuclear_norm", B * N * K * 4):
            _lib.check(self.lib.ogc_mask_nuclear_norm(B, N, K, _ptr(mask), _ptr(out), _ptr(ws), _stream()),
                       "ogc_mask_nuclear_norm")
        self.launches += 2
        return out


_backend = None


def get_backend():
    """The active back-end; the B200 library is loaded on first use and its absence is fatal."""
    global _backend
    if _backend is None:
        _backend = B200Backend()
    return _backend


def set_backend(backend):
    """Install another back-end object (tests: CPU oracle / reference extension). Returns the previous one."""
    global _backend
    prev = _backend
    _backend = backend
    return prev
