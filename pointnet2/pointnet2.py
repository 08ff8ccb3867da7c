"""Drop-in replacement for the reference operator layer `pointnet2/pointnet2.py`
(vLAR-group/OGC), backed by hand-written sm_100a CUDA (libogc_b200.so) instead of the
reference's `pointnet2_cuda` extension.

Same public names, positional signatures, dtypes and return conventions, so that
`from pointnet2.pointnet2 import *` in the reference's utils/pointnet2_util.py:5,
utils/flowstep3d_util.py:4, losses/seg_loss_unsup.py:7 and losses/flow_loss_unsup.py:4 keeps
working unchanged (those star-imports also pick up `torch`, `nn`, `Function`, `Variable`,
`Tuple` from here, hence the imports below).

Reference behaviour, op by op (file:line in /root/reference/pointnet2/pointnet2.py):
  gather_nd :10-14 | furthest_point_sample :17-42 | gather_operation :45-78 | knn :81-109 |
  three_nn :112-140 | three_interpolate :143-187 | grouping_operation :190-230 |
  ball_query :233-260 | QueryAndGroup :263-301 | GroupAll :304-327

All native work goes through `ogc_b200.backend.get_backend()`; there is no CPU fallback in
this module (tests install the CPU oracle through `ogc_b200.backend.set_backend`).
"""
import torch
import torch.nn as nn
from torch.autograd import Function
from torch.autograd import Variable  # noqa: F401  (re-exported: the reference's star-importers see it)
from typing import Tuple  # noqa: F401

from ogc_b200 import backend as _backend_mod


def _be():
    return _backend_mod.get_backend()


def _need_contiguous(**tensors):
    for name, t in tensors.items():
        assert t.is_contiguous(), f"{name} must be contiguous"


def gather_nd(points: torch.Tensor, idx: torch.Tensor, t=False):
    """points (B,N,C) & idx (B,M) int64 -> (B,M,C);  with t=True: points (B,C,N) -> (B,C,M)."""
    if t:
        index = idx[:, None, :].expand(-1, points.shape[1], -1)
        return points.gather(2, index)
    index = idx[:, :, None].expand(-1, -1, points.shape[2])
    return points.gather(1, index)


class FurthestPointSampling(Function):
    """xyz (B,N,3) f32, npoint -> (B,npoint) int32, first index 0, bit-exact to the reference."""

    @staticmethod
    def forward(ctx, xyz: torch.Tensor, npoint: int) -> torch.Tensor:
        _need_contiguous(xyz=xyz)
        sel = _be().fps(xyz, int(npoint))
        ctx.mark_non_differentiable(sel)
        return sel

    @staticmethod
    def backward(ctx, grad=None):
        return None, None


furthest_point_sample = FurthestPointSampling.apply


class GatherOperation(Function):
    """features (B,C,N), idx (B,M) int32 -> (B,C,M); gradient flows to features."""

    @staticmethod
    def forward(ctx, features: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
        _need_contiguous(features=features, idx=idx)
        ctx.n_src = features.shape[2]
        ctx.save_for_backward(idx)
        return _be().gather_points(features, idx)

    @staticmethod
    def backward(ctx, grad_out):
        (idx,) = ctx.saved_tensors
        return _be().gather_points_grad(grad_out.contiguous(), idx, ctx.n_src), None


gather_operation = GatherOperation.apply


class KNN(Function):
    """k, unknown (B,n,3), known (B,m,3) -> (dist (B,n,k) = sqrt(d2), idx (B,n,k) int32),
    neighbours ascending by (distance, index).  Both outputs are fresh, writable tensors
    (callers clip idx in place)."""

    @staticmethod
    def forward(ctx, k: int, unknown: torch.Tensor, known: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        _need_contiguous(unknown=unknown, known=known)
        be = _be()
        if getattr(be, "name", "") == "b200":
            dist, idx = be.knn(int(k), unknown, known, sqrt=True)   # sqrt fused into the kernel's store
        else:
            d2, idx = be.knn(int(k), unknown, known)
            dist = torch.sqrt(d2)
        ctx.mark_non_differentiable(dist, idx)
        return dist, idx

    @staticmethod
    def backward(ctx, a=None, b=None):
        return None, None, None


knn = KNN.apply


class ThreeNN(Function):
    """unknown (B,n,3), known (B,m,3) -> (dist (B,n,3), idx (B,n,3) int32)."""

    @staticmethod
    def forward(ctx, unknown: torch.Tensor, known: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        _need_contiguous(unknown=unknown, known=known)
        d2, idx = _be().three_nn(unknown, known)
        dist = torch.sqrt(d2)
        ctx.mark_non_differentiable(dist, idx)
        return dist, idx

    @staticmethod
    def backward(ctx, a=None, b=None):
        return None, None


three_nn = ThreeNN.apply


class ThreeInterpolate(Function):
    """features (B,C,m), idx (B,n,3) int32, weight (B,n,3) -> (B,C,n); grad to features only."""

    @staticmethod
    def forward(ctx, features: torch.Tensor, idx: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
        _need_contiguous(features=features, idx=idx, weight=weight)
        ctx.m_src = features.shape[2]
        ctx.save_for_backward(idx, weight)
        return _be().three_interpolate(features, idx, weight)

    @staticmethod
    def backward(ctx, grad_out: torch.Tensor):
        idx, weight = ctx.saved_tensors
        return _be().three_interpolate_grad(grad_out.contiguous(), idx, weight, ctx.m_src), None, None


three_interpolate = ThreeInterpolate.apply


class GroupingOperation(Function):
    """features (B,C,N), idx (B,M,S) any integer dtype -> (B,C,M,S); grad to features."""

    @staticmethod
    def forward(ctx, features: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
        _need_contiguous(features=features, idx=idx)
        idx32 = idx.int()
        ctx.n_src = features.shape[2]
        ctx.save_for_backward(idx32)
        return _be().group_points(features, idx32)

    @staticmethod
    def backward(ctx, grad_out: torch.Tensor):
        (idx32,) = ctx.saved_tensors
        return _be().group_points_grad(grad_out.contiguous(), idx32, ctx.n_src), None


grouping_operation = GroupingOperation.apply


class BallQuery(Function):
    """radius, nsample, xyz (B,N,3), new_xyz (B,M,3) -> idx (B,M,nsample) int32."""

    @staticmethod
    def forward(ctx, radius: float, nsample: int, xyz: torch.Tensor, new_xyz: torch.Tensor) -> torch.Tensor:
        _need_contiguous(new_xyz=new_xyz, xyz=xyz)
        idx = _be().ball_query(float(radius), int(nsample), xyz, new_xyz)
        ctx.mark_non_differentiable(idx)
        return idx

    @staticmethod
    def backward(ctx, a=None):
        return None, None, None, None


ball_query = BallQuery.apply


def clip_neighbours_by_radius(dist: torch.Tensor, idx: torch.Tensor, radius) -> torch.Tensor:
    """Neighbours farther than `radius` are replaced by the nearest one (slot 0).  The comparison is
    on the sqrt'ed fp32 distance, as in the reference (:284-286) -- not on squared distances."""
    if radius is None:
        return idx
    return torch.where(dist > radius, idx[..., :1], idx)


class QueryAndGroup(nn.Module):
    """k-NN grouping with radius clipping (the reference's ball_query call is commented out, :281)."""

    def __init__(self, radius: float, nsample: int, use_xyz: bool = True):
        super().__init__()
        self.radius, self.nsample, self.use_xyz = radius, nsample, use_xyz

    def forward(self, xyz: torch.Tensor, new_xyz: torch.Tensor, features: torch.Tensor = None) -> Tuple[torch.Tensor]:
        """xyz (B,N,3), new_xyz (B,M,3), features (B,C,N) ->
        (new_features (B,3+C,M,S) = [xyz - centre, features], grouped_xyz (B,3,M,S))."""
        dist, idx = knn(self.nsample, new_xyz, xyz)
        idx = clip_neighbours_by_radius(dist, idx, self.radius)
        centred = grouping_operation(xyz.transpose(1, 2).contiguous(), idx)
        centred = centred - new_xyz.transpose(1, 2).unsqueeze(-1)
        if features is None:
            assert self.use_xyz, "Cannot have not features and not use xyz as a feature!"
            return centred, centred
        grouped = grouping_operation(features, idx)
        if self.use_xyz:
            grouped = torch.cat([centred, grouped], dim=1)
        return grouped, centred


class GroupAll(nn.Module):
    """Single group holding every point: (B,3+C,1,N)."""

    def __init__(self, use_xyz: bool = True):
        super().__init__()
        self.use_xyz = use_xyz

    def forward(self, xyz: torch.Tensor, new_xyz: torch.Tensor, features: torch.Tensor = None):
        grouped_xyz = xyz.transpose(1, 2).unsqueeze(2)
        if features is None:
            return grouped_xyz, grouped_xyz
        grouped = features.unsqueeze(2)
        if self.use_xyz:
            grouped = torch.cat([grouped_xyz, grouped], dim=1)
        return grouped, grouped_xyz
