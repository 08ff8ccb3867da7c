"""Timing of the two N x N consumers outside the training step at KITTI-SF size (BASELINE.json configs[4] and
SURVEY 8f rank 1): object-aware ICP (B = 64 clouds x 8192 points, K = 10, icp_iter = 20) and multi-frame mask voting
(T = 8 frames x 8192 points, window 3).  Prints one JSON line each.  Not the headline bench (bench.py)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ogc_b200 import backend, icp, vote

be = backend.get_backend()
dev = torch.device("cuda")
torch.manual_seed(0)


def scene(B, N, K):
    pc1 = (torch.rand(B, N, 3, device=dev) - 0.5) * torch.tensor([40.0, 4.0, 30.0], device=dev)
    seg = (pc1[..., 0] / 40.0 + 0.5).mul(K).long().clamp(0, K - 1)
    flow = 0.3 * torch.randn(B, K, 3, device=dev).gather(1, seg.unsqueeze(-1).expand(-1, -1, 3))
    pc2 = (pc1 + flow + 0.01 * torch.randn_like(pc1))[:, torch.randperm(N, device=dev)]
    def soft(s):
        lg = torch.randn(B, N, K, device=dev) * 0.5
        lg.scatter_add_(2, s.unsqueeze(-1), torch.full((B, N, 1), 3.0, device=dev))
        return lg.softmax(-1)
    d = torch.cdist(pc2[:, :, :], (pc1 + flow)[:, :, :]) if N <= 2048 else None
    seg2 = (pc2[..., 0] / 40.0 + 0.5).mul(K).long().clamp(0, K - 1)
    return pc1.contiguous(), pc2.contiguous(), (flow + 0.1 * torch.randn_like(flow)).contiguous(), soft(seg), soft(seg2)


def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        out = fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / reps, out


B, N, K, IT = 64, 8192, 10, 20
pc1, pc2, flow, m1, m2 = scene(B, N, K)
ms, out = timed(lambda: icp.object_aware_icp(pc1, pc2, flow, m1, m2, icp_iter=IT))
print(json.dumps({"workload": f"object_aware_icp B={B} N={N} K={K} icp_iter={IT}", "ms": ms, "clouds_per_s": B / (ms * 1e-3),
                  "pair_evals_per_s": B * N * N * IT / (ms * 1e-3), "finite": bool(torch.isfinite(out).all()),
                  "reference_needs": f"{4 * B * N * N * 4 / 1e9:.0f} GB of (B,N,N) fp32 temporaries per iteration (it runs B=4 chunks)"}))

T = 8
pc = (torch.rand(1, N, 3, device=dev) - 0.5) * torch.tensor([40.0, 4.0, 30.0], device=dev)
pcs = torch.cat([pc + 0.05 * t + 0.01 * torch.randn_like(pc) for t in range(T)], 0).contiguous()
masks = torch.randn(T, N, K, device=dev).mul(2).softmax(-1)
flows = torch.stack([torch.stack([pcs[t + 1] - pcs[t], pcs[t] - pcs[t + 1]]) for t in range(T - 1)]).contiguous()
ms, out = timed(lambda: vote.mask_voting(pcs, masks, flows, time_window_size=3), reps=2)
print(json.dumps({"workload": f"mask_voting T={T} N={N} K={K} window=3", "ms": ms, "frames_per_s": T / (ms * 1e-3),
                  "transfers": 6 * T - 12, "rows_sum_to_one": bool(((out.sum(-1) - 1).abs() < 1e-4).all()),
                  "reference_needs": f"{(T * (T - 1)) * N * N * 4 / 1e9:.0f} GB of (N,N) correspondences and N^3 bmm propagation"}))
