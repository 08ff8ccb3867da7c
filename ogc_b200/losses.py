"""OGC unsupervised segmentation losses -- host-side mirror of the reference's
losses/seg_loss_unsup.py (same public names, constructor arguments and call signatures), in two
implementations:

  * fused   (CUDA tensors + the B200 back-end): the sm_100a kernels of csrc/losses.cu compute each
            loss AND its gradient w.r.t. the soft mask in one pass.  No (B*K,N,N) diag_embed, no
            (B,K,N,S) gathered-mask tensor, no per-sample host syncs (one D2H per step for the
            Hungarian input, one for the logged scalars).
  * composed (any device / back-end): the same arithmetic expressed with torch ops and the
            pointnet2.pointnet2 operator layer, following the reference line by line.  This is what
            runs on the CPU oracle (tests, `bench.py --impl reference`) and what the fused kernels
            are checked against.  It differs from the reference text in one algebraic identity only:
            pc1c^T diag(m) pc2c is evaluated as pc1c^T (m * pc2c) instead of materialising diag_embed(m).

Reference lines are cited per function.
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
from scipy.optimize import linear_sum_assignment
from torch.autograd import Function

import pointnet2.pointnet2 as ops
from ogc_b200 import backend as _backend_mod

FORCE_COMPOSED = False  # tests flip this to compare the two implementations on the same device


# Neighbourhoods computed ahead of time by the trainer (they depend on the coordinates only, so they can run while
# the latency-bound FPS chain occupies a handful of SMs): {(handle, kind, k, radius): (dist, idx)}.  The handle is an
# explicit token the trainer attaches to the very tensor object it later hands to the criterion (`tag_cloud`), not
# the storage address: a recycled allocation can never alias an entry.
NEIGHBOUR_CACHE = {}
_HANDLES = iter(range(1, 1 << 62))


def tag_cloud(pc):
    """Attach a fresh prefetch handle to `pc` (a Python attribute of this tensor object) and return it."""
    pc._ogc_nbr_handle = next(_HANDLES)
    return pc._ogc_nbr_handle


def neighbourhood(be, kind, k, radius, pc):
    """(dist, idx) of one smoothness neighbourhood on the fused path; consumes a prefetched entry when present."""
    handle = getattr(pc, "_ogc_nbr_handle", None)
    hit = NEIGHBOUR_CACHE.pop((handle, kind, k, radius), None) if handle is not None else None
    if hit is not None:
        return hit
    if kind == "knn":
        # neighbours beyond `radius` are replaced by the nearest one anyway (:121-122): bound the search
        return be.knn_bounded(k, pc, pc, radius) if radius is not None else be.knn(k, pc, pc, sqrt=True)
    return None, be.ball_query(radius, k, pc, pc)


def smooth_specs(smooth_loss):
    """The (kind, k, radius) neighbourhoods SmoothLoss will ask for on the fused path (None if it will not fuse)."""
    a, b = smooth_loss.knn_loss, smooth_loss.ball_q_loss
    if a.cross_entropy or b.cross_entropy or a.loss_norm != 1 or b.loss_norm != 1:
        return None
    return (("knn", a.k, a.radius), ("ball", b.k, b.radius))


SIDE_STREAM = True      # run the masks-only latency-bound work of the criterion on a side stream (fused path)
_SIDE = {}


def _side_stream(device):
    key = (device.type, device.index)
    if key not in _SIDE:
        _SIDE[key] = torch.cuda.Stream(device=device)
    return _SIDE[key]


MAX_FUSED_K = 32        # slots the fused loss / matching kernels hold per point (csrc/losses.cu); ICP / transfer: 16
MAX_FUSED_K_ICP = 16


def _use_fused(*tensors, k=None, k_max=MAX_FUSED_K):
    """Fused kernels apply when every tensor is on the GPU, the b200 back-end is active and -- when the number of
    slots `k` is given -- the kernels cover it; otherwise the composed (torch + operator set) path runs, like the
    segnet mask head does (kittidet's n_slot 18 fits the losses but not the ICP kernels)."""
    if FORCE_COMPOSED or not all(t.is_cuda for t in tensors):
        return False
    if k is not None and k > k_max:
        return False
    return getattr(_backend_mod.get_backend(), "name", "") == "b200"


# ------------------------------------------------------------------------------------------------
# Weighted Kabsch                                                    losses/seg_loss_unsup.py:10-61
# ------------------------------------------------------------------------------------------------
def fit_motion_svd_batch(pc1, pc2, mask=None):
    """pc1, pc2 (B,N,3), mask (B,N) or None -> R (B,3,3), t (B,3) minimising sum m |R p1 + t - p2|^2."""
    n_batch = pc1.shape[0]
    if mask is None:
        mu1 = pc1.mean(dim=1, keepdim=True)
        mu2 = pc2.mean(dim=1, keepdim=True)
    else:
        w = mask.sum(dim=1, keepdim=True)
        mu1 = (torch.einsum("bnd,bn->bd", pc1, mask) / w).unsqueeze(1)
        mu2 = (torch.einsum("bnd,bn->bd", pc2, mask) / w).unsqueeze(1)
    c1, c2 = pc1 - mu1, pc2 - mu2
    if mask is not None:
        c2 = c2 * mask.unsqueeze(-1)          # == diag_embed(mask) @ c2 without the (B,N,N) tensor (:36)
    S = torch.bmm(c1.transpose(1, 2), c2)

    ok = ~torch.isnan(S).flatten(1).any(dim=1)                      # ill-posed segments -> identity (:40-42)
    R_all = torch.eye(3, device=pc1.device, dtype=pc1.dtype).repeat(n_batch, 1, 1)
    t_all = torch.zeros(n_batch, 3, device=pc1.device, dtype=pc1.dtype)
    if ok.any():
        U, _, Vh = torch.linalg.svd(S[ok])
        V = Vh.transpose(1, 2)
        d = torch.ones_like(S[ok][..., 0])
        d[:, 2] = torch.det(torch.bmm(V, U.transpose(1, 2)))         # reflection -> rotation (:47-53)
        R = torch.bmm(V * d.unsqueeze(1), U.transpose(1, 2))
        t = mu2[ok].squeeze(1) - torch.bmm(R, mu1[ok].transpose(1, 2)).squeeze(2)
        R_all[ok] = R
        t_all[ok] = t
    return R_all, t_all


def _per_object_rigid(pc, pc2, mask):
    """Composed helper: per-object transforms of `pc` for soft mask (B,N,K) -> (B,K,N,3)."""
    B, N, K = mask.shape
    m = mask.transpose(1, 2).reshape(B * K, N)
    rep = lambda x: x.unsqueeze(1).expand(B, K, N, 3).reshape(B * K, N, 3)
    R, t = fit_motion_svd_batch(rep(pc), rep(pc2), m)
    moved = torch.einsum("bij,bnj->bni", R, rep(pc)) + t.unsqueeze(1)
    return moved.reshape(B, K, N, 3)


class _DynamicLossFn(Function):
    @staticmethod
    def forward(ctx, pc, mask, flow):
        be = _backend_mod.get_backend()
        loss_pt, grad, _ = be.dynamic_loss(pc.contiguous(), flow.contiguous(), mask.contiguous(),
                                           need_grad=mask.requires_grad)
        ctx.count = loss_pt.numel()
        ctx.save_for_backward(grad)
        return loss_pt.mean()

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return None, grad * (g / ctx.count), None


class DynamicLoss(nn.Module):
    """mean || sum_k m_k (R_k p + t_k) - (p + flow) ||; R,t from weighted Kabsch, treated as constants
    in the backward pass (losses/seg_loss_unsup.py:64-98)."""

    def __init__(self, loss_norm=2):
        super().__init__()
        self.loss_norm = loss_norm

    def forward(self, pc, mask, flow):
        if self.loss_norm == 2 and _use_fused(pc, mask, flow, k=mask.shape[-1]):
            return _DynamicLossFn.apply(pc, mask, flow)
        pc2 = pc + flow
        moved = _per_object_rigid(pc, pc2, mask).detach()                       # (B,K,N,3)
        blended = (mask.transpose(1, 2).unsqueeze(-1) * moved).sum(dim=1)
        return (blended - pc2).norm(p=self.loss_norm, dim=-1).mean()


# ------------------------------------------------------------------------------------------------
# Smoothness                                                        losses/seg_loss_unsup.py:101-180
# ------------------------------------------------------------------------------------------------
class _NeighborL1Fn(Function):
    """sum_i coef_i * mean_{b,n} mean_s |m_n - m_nbr|_1 for a list of neighbourhoods (fused fwd+bwd)."""

    @staticmethod
    def forward(ctx, mask, pc, specs):
        be = _backend_mod.get_backend()
        mask = mask.contiguous()
        grad = torch.zeros_like(mask) if mask.requires_grad else None
        total = None
        for kind, k, radius, coef in specs:
            dist, idx = neighbourhood(be, kind, k, radius, pc)
            if kind == "knn":
                loss_pt = be.neighbor_l1(mask, idx, dist if radius is not None else None, radius, coef, grad)
            else:
                loss_pt = be.neighbor_l1(mask, idx, None, None, coef, grad)
            term = loss_pt.mean() * coef
            total = term if total is None else total + term
        ctx.count = mask.shape[0] * mask.shape[1]
        ctx.save_for_backward(grad)
        return total

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return grad * (g / ctx.count), None, None


def _neighbor_loss_composed(mask, idx, cross_entropy, loss_norm):
    m = mask.permute(0, 2, 1).contiguous()                                     # (B,K,N)
    nn_mask = ops.grouping_operation(m, idx.detach())                          # (B,K,N,S)
    if cross_entropy:
        tgt = m.unsqueeze(3).expand_as(nn_mask).detach()
        loss = F.binary_cross_entropy(nn_mask, tgt, reduction="none").sum(dim=1).mean(dim=-1)
    else:
        loss = (m.unsqueeze(3) - nn_mask).norm(p=loss_norm, dim=1).mean(dim=-1)
    return loss.mean()


class KnnLoss(nn.Module):
    """:101-129  k-NN neighbourhood, neighbours beyond `radius` replaced by the nearest one."""

    def __init__(self, k, radius, cross_entropy=False, loss_norm=1, **kwargs):
        super().__init__()
        self.k, self.radius, self.cross_entropy, self.loss_norm = k, radius, cross_entropy, loss_norm

    def forward(self, pc, mask):
        if not self.cross_entropy and self.loss_norm == 1 and _use_fused(pc, mask, k=mask.shape[-1]):
            return _NeighborL1Fn.apply(mask, pc.contiguous(), (("knn", self.k, self.radius, 1.0),))
        dist, idx = ops.knn(self.k, pc, pc)
        idx = ops.clip_neighbours_by_radius(dist, idx, self.radius)
        return _neighbor_loss_composed(mask, idx, self.cross_entropy, self.loss_norm)


class BallQLoss(nn.Module):
    """:132-158  ball-query neighbourhood."""

    def __init__(self, k, radius, cross_entropy=False, loss_norm=1, **kwargs):
        super().__init__()
        self.k, self.radius, self.cross_entropy, self.loss_norm = k, radius, cross_entropy, loss_norm

    def forward(self, pc, mask):
        if not self.cross_entropy and self.loss_norm == 1 and _use_fused(pc, mask, k=mask.shape[-1]):
            return _NeighborL1Fn.apply(mask, pc.contiguous(), (("ball", self.k, self.radius, 1.0),))
        idx = ops.ball_query(self.radius, self.k, pc, pc)
        return _neighbor_loss_composed(mask, idx, self.cross_entropy, self.loss_norm)


class SmoothLoss(nn.Module):
    """:161-180  w_knn * KnnLoss + w_ball_q * BallQLoss."""

    def __init__(self, w_knn, w_ball_q, knn_loss_params, ball_q_loss_params):
        super().__init__()
        self.knn_loss = KnnLoss(**knn_loss_params)
        self.ball_q_loss = BallQLoss(**ball_q_loss_params)
        self.w_knn, self.w_ball_q = w_knn, w_ball_q

    def forward(self, pc, mask):
        a, b = self.knn_loss, self.ball_q_loss
        plain = not (a.cross_entropy or b.cross_entropy) and a.loss_norm == 1 and b.loss_norm == 1
        if plain and _use_fused(pc, mask, k=mask.shape[-1]):
            return _NeighborL1Fn.apply(mask, pc.contiguous(), (("knn", a.k, a.radius, float(self.w_knn)),
                                                               ("ball", b.k, b.radius, float(self.w_ball_q))))
        return self.w_knn * a(pc, mask) + self.w_ball_q * b(pc, mask)


# ------------------------------------------------------------------------------------------------
# Invariance                                                        losses/seg_loss_unsup.py:183-280
# ------------------------------------------------------------------------------------------------
def interpolate_mask_by_flow(pc1, pc2, mask1, flow1, k=1):
    """Mask of pc2 from the k nearest warped points of pc1 (:183-209).  -> (B,N,K)"""
    warped = (pc1 + flow1).contiguous()
    dist, idx = ops.knn(k, pc2.contiguous(), warped)
    picked = ops.grouping_operation(mask1.transpose(1, 2).contiguous(), idx.detach())      # (B,K,N,k)
    if k == 1:
        out = picked.squeeze(-1)
    else:
        inv = 1.0 / dist.clamp(min=1e-10)
        w = inv / inv.sum(dim=2, keepdim=True)
        out = (w.unsqueeze(1) * picked).sum(dim=-1)
    return out.transpose(1, 2)


def _hungarian_from_counts(inter):
    """inter (B,K,K) integer contingency (numpy) -> col_ind (B,K) maximising the fp32 IoU, exactly as
    :226-237 computes it: intersection / clamp(|A| + |B| - intersection, 1e-10), scipy LSA per sample."""
    inter = inter.astype(np.float32)
    union = inter.sum(axis=2, keepdims=True) + inter.sum(axis=1, keepdims=True) - inter
    iou = inter / np.maximum(union, np.float32(1e-10))
    return np.stack([linear_sum_assignment(iou[b], maximize=True)[1] for b in range(iou.shape[0])], 0)


def match_indices_by_iou(mask1, mask2):
    """-> (perm12, perm21) (B,K): slot of mask2 matched to each slot of mask1, and the converse
    (= match_mask_by_iou(mask1, mask2) and match_mask_by_iou(mask2, mask1) of the reference).
    Fused path: int32 CUDA tensors computed entirely on the device (contingency kernel + device Hungarian, no
    host sync); composed path: int64 numpy arrays through scipy, as the reference."""
    if _use_fused(mask1, mask2, k=mask1.shape[-1]):
        be = _backend_mod.get_backend()
        inter = be.mask_contingency(mask1.detach().contiguous(), mask2.detach().contiguous())
        return be.mask_match(inter)
    K = mask1.shape[2]
    a1, a2 = mask1.argmax(-1), mask2.argmax(-1)
    inter = torch.zeros(mask1.shape[0], K * K, dtype=torch.int64, device=mask1.device)
    inter.scatter_add_(1, a1 * K + a2, torch.ones_like(a1))
    inter = inter.view(-1, K, K).cpu().numpy()
    return _hungarian_from_counts(inter), _hungarian_from_counts(inter.transpose(0, 2, 1))


def match_mask_by_iou(mask1, mask2):
    """(B,N,K) x2 -> permutation matrices (B,K,K) aligning mask2's slots to mask1's (:212-240)."""
    perm12, _ = match_indices_by_iou(mask1, mask2)
    if not torch.is_tensor(perm12):
        perm12 = torch.from_numpy(perm12).to(mask1.device)
    eye = torch.eye(mask1.shape[2], dtype=torch.float32, device=mask1.device)
    return eye[perm12.long()]


class _InvarianceFn(Function):
    @staticmethod
    def forward(ctx, mask1, mask2, perm12, perm21):
        be = _backend_mod.get_backend()
        need = mask1.requires_grad or mask2.requires_grad
        loss_pt, g1, g2 = be.invariance_loss(mask1.contiguous(), mask2.contiguous(), perm12, perm21, need_grad=need)
        ctx.count = loss_pt.numel()
        ctx.save_for_backward(g1, g2)
        return loss_pt.mean()

    @staticmethod
    def backward(ctx, g):
        g1, g2 = ctx.saved_tensors
        s = g / ctx.count
        return g1 * s, g2 * s, None, None


class InvarianceLoss(nn.Module):
    """:243-280  distance between each view's mask and the other view's mask permuted by the Hungarian match."""

    def __init__(self, cross_entropy=False, loss_norm=2):
        super().__init__()
        self.cross_entropy, self.loss_norm = cross_entropy, loss_norm

    def distance(self, pred, target):
        if self.cross_entropy:
            return F.binary_cross_entropy(pred, target, reduction="none").sum(dim=1).mean()
        return (pred - target).norm(p=self.loss_norm, dim=-1).mean()

    def forward(self, mask1, mask2, perms=None):
        """`perms`: (perm12, perm21) when the matching was computed ahead of time (side stream)."""
        perm12, perm21 = perms if perms is not None else match_indices_by_iou(mask1, mask2)
        dev = mask1.device
        if torch.is_tensor(perm12):                                   # fused matching: already on the device
            if not self.cross_entropy and self.loss_norm == 2:
                return _InvarianceFn.apply(mask1, mask2, perm12, perm21)
            perm12, perm21 = perm12.long(), perm21.long()
        else:
            perm12, perm21 = torch.from_numpy(perm12).to(dev), torch.from_numpy(perm21).to(dev)
        i12 = perm12.unsqueeze(1).expand_as(mask1)
        i21 = perm21.unsqueeze(1).expand_as(mask2)
        target1 = torch.gather(mask2, 2, i12).detach()             # == einsum('bij,bnj->bni', perm2, mask2)
        target2 = torch.gather(mask1, 2, i21).detach()
        return self.distance(mask1, target1) + self.distance(mask2, target2)


# ------------------------------------------------------------------------------------------------
# Monitoring terms + the combined loss                              losses/seg_loss_unsup.py:283-409
# ------------------------------------------------------------------------------------------------
class EntropyLoss(nn.Module):
    def forward(self, mask, epsilon=1e-5):
        return -(mask * torch.log(mask.clamp(epsilon))).sum(dim=-1).mean()


class RankLoss(nn.Module):
    """mean nuclear norm of the (N,K) masks (:300-314).  On the fused path the singular values come from
    the K x K Gram matrix accumulated in fp64 (sigma = sqrt(eig(M^T M)), Jacobi on the device) instead of an
    (N,K) cuSOLVER SVD with its host sync."""

    def forward(self, mask):
        # the fused kernel is forward-only (the reference logs this term and never back-propagates it): a mask that
        # requires grad outside torch.no_grad() keeps the differentiable composed expression
        if _use_fused(mask, k=mask.shape[-1]) and not (mask.requires_grad and torch.is_grad_enabled()):
            return _backend_mod.get_backend().mask_nuclear_norm(mask.detach().contiguous()).mean()
        return mask.norm(p="nuc", dim=(1, 2)).mean()


class UnsupervisedOGCLoss(nn.Module):
    """loss = w_dyn * dynamic + w_smooth * smooth (+ w_inv * invariance with augmentation); entropy and
    rank are logged only.  With `aug_transform` the four views contribute with the reference's 0.5
    factor (:358-405)."""

    def __init__(self, dynamic_loss, smooth_loss, invariance_loss, entropy_loss, rank_loss,
                 weights=[10.0, 0.1, 0.1], start_steps=[0, 0, 0]):
        super().__init__()
        self.dynamic_loss, self.smooth_loss, self.invariance_loss = dynamic_loss, smooth_loss, invariance_loss
        self.entropy_loss, self.rank_loss = entropy_loss, rank_loss
        self.w_dynamic, self.w_smooth, self.w_invariance = weights
        self.start_step_dynamic, self.start_step_smooth, self.start_step_invariance = start_steps
        # True: loss_dict = {"_keys": [...], "_values": device tensor} -- no host sync inside forward (needed to
        # capture the whole step in a CUDA graph); resolve_loss_dict() turns it into the reference's dict of floats.
        self.defer_logging = False

    def step_lossw(self, it, weight, start_step=0):
        return 0 if it < start_step else weight

    def forward(self, pcs, masks, flows, step_w=False, it=0, aug_transform=False):
        assert len(pcs) == len(masks) == len(flows), "Inconsistent number of frames!"
        n_view = 4 if aug_transform else 2
        assert len(pcs) == n_view
        scale = 0.5 if aug_transform else 1.0
        w = lambda weight, start: self.step_lossw(it, weight, start) if step_w else weight

        # Latency-bound work that depends on the masks only -- the device Hungarian matching of the invariance term
        # (one warp per sample) and the logged-only entropy / nuclear norm (Jacobi on K x K Gram matrices) -- runs on
        # a side stream underneath the dynamic / smoothness kernels (fork / join by events: parallel branches of the step graph).
        fused_all = _use_fused(*masks, k=masks[0].shape[-1]) and len({m.shape for m in masks}) == 1
        side = main = None
        perms = [None, None]
        logged = {}

        def logged_terms():
            with torch.no_grad():
                if fused_all:
                    # logged-only terms: one launch over all views (sum of per-view means = n_view * mean over all)
                    allm = torch.cat([m.detach() for m in masks], dim=0)
                    logged["entropy"] = scale * n_view * self.entropy_loss(allm)
                    logged["rank"] = scale * n_view * self.rank_loss(allm)
                else:
                    logged["entropy"] = scale * sum(self.entropy_loss(m) for m in masks)
                    logged["rank"] = scale * sum(self.rank_loss(m) for m in masks)

        if fused_all and SIDE_STREAM:
            main = torch.cuda.current_stream()
            side = _side_stream(masks[0].device)
            side.wait_stream(main)
            with torch.cuda.stream(side):
                if aug_transform:
                    perms = [match_indices_by_iou(masks[0], masks[2]), match_indices_by_iou(masks[1], masks[3])]
                logged_terms()
        if _use_fused(*pcs, *masks, *flows, k=masks[0].shape[-1]) and len({m.shape for m in masks}) == 1:
            # the Kabsch kernel runs one CTA per cloud: all views in ONE launch (sum of per-view means = n_view * mean)
            l_dynamic = scale * n_view * self.dynamic_loss(torch.cat(pcs, 0), torch.cat(masks, 0), torch.cat(flows, 0))
        else:
            l_dynamic = scale * sum(self.dynamic_loss(pcs[v], masks[v], flows[v]) for v in range(n_view))
        l_smooth = scale * sum(self.smooth_loss(pcs[v], masks[v]) for v in range(n_view))
        if side is not None:
            main.wait_stream(side)
        terms = [w(self.w_dynamic, self.start_step_dynamic) * l_dynamic,
                 w(self.w_smooth, self.start_step_smooth) * l_smooth]
        ordered = {"dynamic": l_dynamic, "smooth": l_smooth}
        if aug_transform:
            l_inv = self.invariance_loss(masks[0], masks[2], perms[0]) + self.invariance_loss(masks[1], masks[3], perms[1])
            terms.append(w(self.w_invariance, self.start_step_invariance) * l_inv)
            ordered["invariance"] = l_inv
        if side is None:
            logged_terms()
        ordered["entropy"], ordered["rank"] = logged["entropy"], logged["rank"]
        logged = ordered
        loss = sum(terms)
        logged["sum"] = loss
        # one device->host transfer for every logged scalar (the reference calls .item() six times)
        keys = list(logged)
        if self.defer_logging:
            return loss, {"_keys": keys, "_values": torch.stack([logged[k].detach().float().reshape(()) for k in keys])}
        values = torch.stack([logged[k].detach().float().reshape(()) for k in keys]).tolist()
        loss_dict = dict(zip(keys, values))
        loss_dict.setdefault("invariance", 0)
        return loss, loss_dict


def resolve_loss_dict(d, values=None):
    """Deferred loss dict -> the reference's dict of Python floats (one device->host read)."""
    if "_keys" not in d:
        return d
    vals = (d["_values"] if values is None else values).tolist()
    out = dict(zip(d["_keys"], vals))
    out.setdefault("invariance", 0)
    return out


def build_ogc_loss(loss_cfg):
    """Construct the criterion from the `loss:` block of a reference yaml (train_seg.py:332-345)."""
    return UnsupervisedOGCLoss(
        DynamicLoss(**loss_cfg["dynamic_loss_params"]), SmoothLoss(**loss_cfg["smooth_loss_params"]),
        InvarianceLoss(**loss_cfg["invariance_loss_params"]), EntropyLoss(), RankLoss(),
        weights=loss_cfg["weights"], start_steps=loss_cfg["start_steps"])


KITTISF_LOSS_CFG = {   # config/seg/kittisf/kittisf_unsup.yaml:39-56
    "weights": [10.0, 0.1, 0.1], "start_steps": [0, 100, 1000],
    "dynamic_loss_params": {"loss_norm": 2},
    "smooth_loss_params": {"w_knn": 3.0, "w_ball_q": 1.0,
                           "knn_loss_params": {"k": 32, "radius": 1.0, "loss_norm": 1},
                           "ball_q_loss_params": {"k": 64, "radius": 2.0, "loss_norm": 1}},
    "invariance_loss_params": {"loss_norm": 2},
}
