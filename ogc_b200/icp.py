"""Object-aware ICP -- host-side mirror of the functions of the reference's oa_icp.py (:16-84), same names and
signatures: `weighted_kabsch(pc, flow, mask)`, `object_aware_icp(pc1, pc2, flow, mask1, mask2, icp_iter, temperature)`.

Fused path (CUDA + B200 back-end): per iteration ONE correspondence kernel (online softmax over the target cloud,
no (B,N,N) tensor: csrc/icp.cu) + the per-object weighted Kabsch kernel (csrc/losses.cu) + the rigid blend kernel;
the slot matching uses the device Hungarian.  Composed path: the reference's formulation with torch ops (CPU oracle
tests, float64 evaluation).
"""
import torch

import pointnet2.pointnet2 as ops
from ogc_b200 import backend as _backend_mod
from ogc_b200.losses import MAX_FUSED_K_ICP, _use_fused, fit_motion_svd_batch, interpolate_mask_by_flow, match_mask_by_iou, \
    match_indices_by_iou


def weighted_kabsch(pc, flow, mask):
    """pc, flow (B,N,3), mask (B,N,K) -> flow (B,N,3) projected on per-object rigid motions (oa_icp.py:16-38)."""
    if _use_fused(pc, flow, mask, k=mask.shape[-1]):
        be = _backend_mod.get_backend()
        pc, flow, mask = pc.contiguous(), flow.contiguous(), mask.contiguous()
        Rt = be.weighted_kabsch(pc, flow, mask, second_is_flow=True)
        return be.apply_rigid_flow(pc, mask, Rt)
    B, N, K = mask.shape
    m = mask.transpose(1, 2).reshape(B * K, N)
    rep = lambda x: x.unsqueeze(1).expand(B, K, N, 3).reshape(B * K, N, 3)
    R, t = fit_motion_svd_batch(rep(pc), rep(pc + flow), m)
    moved = (torch.einsum("bij,bnj->bni", R, rep(pc)) + t.unsqueeze(1)).reshape(B, K, N, 3)
    return torch.einsum("bkn,bkni->bni", m.reshape(B, K, N), moved) - pc


def object_aware_icp(pc1, pc2, flow, mask1, mask2, icp_iter=10, temperature=0.01):
    """oa_icp.py:41-84.  pc1, pc2, flow (B,N,3), mask1, mask2 (B,N,K) -> refined flow (B,N,3)."""
    fused = _use_fused(pc1, pc2, flow, mask1, mask2, k=mask1.shape[-1], k_max=MAX_FUSED_K_ICP)
    # align the slot order of frame 2 to frame 1 (:52-54)
    mask2_interp = interpolate_mask_by_flow(pc1, pc2, mask1, flow)
    if fused:
        perm, _ = match_indices_by_iou(mask2_interp.contiguous(), mask2)
        mask2 = torch.gather(mask2, 2, perm.long().unsqueeze(1).expand_as(mask2)).contiguous()
        be = _backend_mod.get_backend()
        pc1, pc2, mask1 = pc1.contiguous(), pc2.contiguous(), mask1.contiguous()
        flow = flow.contiguous()
        for _ in range(icp_iter):
            flow = be.icp_correspond(pc1, flow, pc2, mask1, mask2, temperature)          # (:66-75)
            Rt = be.weighted_kabsch(pc1, flow, mask1, second_is_flow=True)                # (:77-79)
            flow = be.apply_rigid_flow(pc1, mask1, Rt)                                    # (:80-83)
        return flow
    perm = match_mask_by_iou(mask2_interp, mask2).to(mask2.dtype)
    mask2 = torch.einsum("bij,bnj->bni", perm, mask2)
    consistency = torch.einsum("bmk,bnk->bmn", mask1, mask2)
    for _ in range(icp_iter):
        corr = (-torch.cdist(pc1 + flow, pc2) / temperature).softmax(-1) * consistency
        corr = corr / corr.sum(-1, keepdim=True).clamp(1e-10)
        flow = torch.einsum("bmn,bnj->bmj", corr, pc2) - pc1
        flow = weighted_kabsch(pc1, flow, mask1)
    return flow
