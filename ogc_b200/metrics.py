"""On-device segmentation metrics -- mirror of `accumulate_eval_results` / `eval_segm` (metrics/seg_metric.py:8-93),
same name, arguments and return values.

The reference moves the ground-truth labels and the soft masks to the host and loops over samples, GT objects and
predicted objects in numpy (n_gt x n_pred boolean reductions over N points each).  Here the whole batch is evaluated
with a handful of device-wide tensor operations (one flat bincount gives every sample's GT x prediction contingency
table) and ONE device->host transfer of (B, K) result rows; no per-sample Python loop touches point data.
"""
import numpy as np
import torch


def accumulate_eval_results(segm, mask, ignore_npoint_thresh=0):
    """segm (B,N) integer GT labels, mask (B,N,K) soft masks -> (Pred_IoU, Pred_Matched, Confidence, N_GT_Inst):
    three 1-D numpy arrays over all VALID predicted objects of the batch (sample-major, slot ascending) and an int."""
    B, N, K = mask.shape
    dev = mask.device
    segm = segm.to(dev).long()
    segm = segm - segm.min()                                   # labels are arbitrary ints (np.unique relabels them)
    G = int(segm.max().item()) + 1
    mx = mask.max(dim=2, keepdim=True).values
    ar = torch.arange(K, device=dev).view(1, 1, K)
    pred = torch.where(mask == mx, ar, K).min(dim=2).values    # first maximum, like np.argmax (:49)
    base = torch.arange(B, device=dev).view(B, 1)
    inter = torch.bincount(((base * G + segm) * K + pred).reshape(-1), minlength=B * G * K).view(B, G, K).double()
    gt_sizes = inter.sum(2)                                    # (B,G)   0 = label absent in this sample
    pred_sizes = inter.sum(1)                                  # (B,K)
    present = gt_sizes > 0
    ignored = present & (gt_sizes < ignore_npoint_thresh)      # too small GT objects (:60)
    ign_area = (inter * ignored.unsqueeze(2)).sum(1)           # (B,K)
    safe = pred_sizes.clamp_min(1)
    invalid = (ign_area / safe) > 0.5                          # an FP mostly on ignored GT is not penalised (:63-64)
    adj_sizes = pred_sizes - ign_area                          # (:67)
    valid = (pred_sizes > 0) & (adj_sizes > 0) & ~invalid      # (:68)
    kept = present & ~ignored
    # soft-mask mass of every slot c over the points ASSIGNED to slot s: T[b,s,c] = sum_{pred == s} mask[:, c]
    onehot = torch.zeros(B, N, K, dtype=mask.dtype, device=dev).scatter_(2, pred.unsqueeze(2), 1.0)
    T = torch.einsum("bns,bnc->bsc", onehot.double(), mask.double())
    union = gt_sizes.unsqueeze(2) + adj_sizes.unsqueeze(1) - inter
    iou = torch.where(kept.unsqueeze(2), inter / union.clamp_min(1e-300), torch.full_like(inter, -1.0))
    pred_iou = iou.max(dim=1).values                           # (B,K) (:90)
    packed = torch.cat([pred_iou, valid.double(), pred_sizes, T.reshape(B, K * K)], 1)
    packed_h, n_gt = packed.cpu().numpy(), int(kept.sum().item())      # the single device->host read of the results
    Pred_IoU, Confidence = [], []
    for b in range(B):                                         # K-sized bookkeeping only, no point data
        iou_b, valid_b, sizes_b = packed_h[b, :K], packed_h[b, K:2 * K] > 0.5, packed_h[b, 2 * K:3 * K]
        T_b = packed_h[b, 3 * K:].reshape(K, K)
        present_slots = np.nonzero(sizes_b > 0)[0]             # np.unique(segm_pred) (:51)
        valid_slots = np.nonzero(valid_b)[0]
        Pred_IoU.append(iou_b[valid_slots])
        # Reference quirk kept (seg_metric.py:73-85): after invalid predictions are dropped the mask COLUMNS are
        # re-indexed but the per-point instance ids are not, so confidence[j] averages column valid_slots[j] over the
        # points of the j-th PRESENT prediction.  Identical to the intended value whenever nothing is dropped.
        Confidence.append(np.array([T_b[present_slots[j], valid_slots[j]] / sizes_b[present_slots[j]]
                                    for j in range(len(valid_slots))], dtype=np.float64))
    Pred_IoU = np.concatenate(Pred_IoU) if Pred_IoU else np.zeros(0)
    Confidence = np.concatenate(Confidence) if Confidence else np.zeros(0)
    return Pred_IoU, (Pred_IoU >= 0.5).astype(float), Confidence, n_gt
