"""Multi-frame mask voting -- host-side mirror of the reference's vote.py (:17-131), same function names and
signatures: `pairwise_correspondence`, `match_mask_by_cost`, `mask_voting`.

The reference materialises every pairwise soft correspondence as an (N,N) matrix (268 MB per pair at N = 8192) and
propagates non-adjacent ones with N x N x N `bmm`s (vote.py:50-57).  All the voting ever does with a correspondence is
`corr @ mask` (vote.py:121), and a product of row-stochastic matrices is row-stochastic, so

    corr(t, t+2) @ M  =  corr(t, t+1) @ (corr(t+1, t+2) @ M)

and each factor is one streaming softmax-weighted average (csrc/icp.cu `softmax_transfer_kernel`, O(N^2 K) work,
no N x N tensor, chains shared between target frames).  The composed path evaluates the same chain with dense torch ops
(CPU / float64 tests).
"""
import torch
import torch.nn.functional as F
from scipy.optimize import linear_sum_assignment

from ogc_b200 import backend as _backend_mod
from ogc_b200.losses import _use_fused


def pairwise_correspondence(pc1, pc2, flow, temperature=0.01):
    """vote.py:17-28 -- dense (B,N,N) soft correspondence; kept for API parity and tests (the voting does not use it)."""
    return (-torch.cdist(pc1 + flow, pc2, compute_mode="donot_use_mm_for_euclid_dist") / temperature).softmax(-1)


def transfer(src_warped, dst, values, temperature=0.01):
    """softmax(-cdist(src_warped, dst) / T) @ values without the matrix.  (B,N1,3), (B,N2,3), (B,N2,K) -> (B,N1,K)."""
    if _use_fused(src_warped, dst, values) and values.shape[-1] <= 16:
        return _backend_mod.get_backend().softmax_transfer(src_warped.contiguous(), dst.contiguous(),
                                                           values.contiguous(), temperature)
    return torch.bmm(pairwise_correspondence(src_warped, dst, torch.zeros_like(src_warped), temperature), values)


def match_mask_by_cost(mask1, mask2, measure="ce"):
    """vote.py:62-92 -- reorder the slots of mask2 (N,K) to those of mask1 (N,K) with the Hungarian algorithm on the
    mean binary cross-entropy ('ce') or the soft IoU.  K x K problem on the host, as the reference."""
    n_object = mask1.shape[-1]
    m1 = mask1.unsqueeze(2).expand(-1, -1, n_object)
    m2 = mask2.unsqueeze(1).expand(-1, n_object, -1)
    if measure == "ce":
        cost = F.binary_cross_entropy(m1, m2, reduction="none").mean(0)
        _, col_ind = linear_sum_assignment(cost.detach().cpu().numpy(), maximize=False)
    else:
        iou = (m1 * m2).sum(0) / (m1 + m2).sum(0).clamp(1e-10)
        _, col_ind = linear_sum_assignment(iou.detach().cpu().numpy(), maximize=True)
    return mask2[:, torch.as_tensor(col_ind, device=mask2.device)]      # == einsum('ij,nj->ni', eye[col_ind], mask2)


def mask_voting(pc, mask, flows, time_window_size=3, temperature=0.01):
    """vote.py:95-131.  pc (T,N,3), mask (T,N,K), flows (T-1,2,N,3) adjacent forward / backward flows -> (T,N,K)."""
    n_frame = pc.shape[0]
    votes = [[None] * n_frame for _ in range(n_frame)]        # votes[t][v] = corr(t, v) @ mask[v]
    for v in range(n_frame):
        # frames before v: corr(t,v) = corr(t,t+1) ... corr(v-1,v), applied right to left
        x = mask[v:v + 1]
        for t in range(v - 1, max(v - time_window_size, 0) - 1, -1):
            x = transfer(pc[t:t + 1] + flows[t:t + 1, 0], pc[t + 1:t + 2], x, temperature)
            votes[t][v] = x[0]
        # frames after v: corr(t,v) = corr(t,t-1) ... corr(v+1,v)
        x = mask[v:v + 1]
        for t in range(v + 1, min(v + time_window_size, n_frame - 1) + 1):
            x = transfer(pc[t:t + 1] + flows[t - 1:t, 1], pc[t - 1:t], x, temperature)
            votes[t][v] = x[0]
    voted = []
    for t in range(n_frame):
        window = range(max(0, t - time_window_size), min(n_frame, t + time_window_size + 1))
        stack = [mask[t] if v == t else match_mask_by_cost(mask[t], votes[t][v]) for v in window]
        vote = torch.stack(stack, 0).mean(0)
        voted.append(vote / vote.sum(-1, keepdim=True).clamp(1e-10))
    return torch.stack(voted, 0)
