"""Per-launch timing of the SA MLP kernels at KITTI-SF sizes (B=16): chained (round 2) vs per-layer (round 1).

    gpurun -- python scratch/chain_bench.py [reps]      # CHAIN=0 for the round-1 kernels, CFGS=SA2,SA3 to select
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ogc_b200 import backend, segnet, sa_fused
from ogc_b200.sa_fused import fused_sa_mlp
import pointnet2.pointnet2 as ops

be = backend.get_backend()
B = 16
torch.manual_seed(0)
cfgs = [("SA1a", 8192, 2048, 3, [32, 32, 32]), ("SA1b", 8192, 2048, 3, [32, 32, 64]), ("SA2", 2048, 1024, 96, [64, 64, 128]),
        ("SA3", 1024, 512, 128, [128, 128, 256])]
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
if os.environ.get("CFGS"):
    cfgs = [c for c in cfgs if c[0] in os.environ["CFGS"].split(",")]
sa_fused.USE_CHAIN = os.environ.get("CHAIN", "1") == "1"       # default here: the chained kernels
sa_fused.STORE_Y = os.environ.get("STORE_Y", "1") == "1"
sa_fused.USE_CHAIN_DX = {"0": False, "1": True}.get(os.environ.get("CHAIN_DX", "1"), os.environ.get("CHAIN_DX"))
sa_fused.USE_DW_TMA = os.environ.get("DW_TMA", "1") == "1"
sa_fused.DW_TMA_NARROW = os.environ.get("DW_TMA_NARROW", "0") == "1"
sa_fused.USE_FWD_TMA = os.environ.get("FWD_TMA", "1") == "1"
sa_fused.FWD_TMA_NARROW = os.environ.get("FWD_TMA_NARROW", "1") == "1"
sa_fused.USE_DX_TMA = {"0": False, "1": True}.get(os.environ.get("DX_TMA", "auto"), os.environ.get("DX_TMA", "auto"))
sa_fused.USE_NARROW = os.environ.get("NARROW", "1") == "1"
fwd_only = os.environ.get("FWD_ONLY", "0") == "1"
for name, N, M, Cf, w in cfgs:
    xyz = (torch.rand(B, N, 3, device="cuda") - 0.5) * 40
    new_xyz = xyz[:, :M].contiguous()
    feat = torch.randn(B, N, Cf, device="cuda", requires_grad=True)
    mlp = segnet.SharedMLP([Cf + 3] + w).cuda()
    dist, idx = ops.knn(64, new_xyz, xyz)
    layers = [(getattr(mlp, f"layer{i}").conv.weight, getattr(mlp, f"layer{i}").normlayer.gn.weight,
               getattr(mlp, f"layer{i}").normlayer.gn.bias) for i in range(3)]
    probe = torch.randn(B, w[-1], M, device="cuda")
    for r in range(reps + 1):
        if r == 1:
            torch.cuda.synchronize(); backend.TIMER.enabled = True; backend.TIMER.detail = True; backend.TIMER.reset()
        out = fused_sa_mlp(xyz, new_xyz, feat, idx, layers)
        if not fwd_only:
            (out * probe).sum().backward()
        if os.environ.get("VERBOSE"):
            torch.cuda.synchronize(); print("rep", r, "done", flush=True)
    torch.cuda.synchronize(); backend.TIMER.enabled = False
    P = M * 64
    print(f"== {name}: N={N} M={M} P={P} Cin={Cf + 3} widths={w} chain={sa_fused.USE_CHAIN}")
    tot = 0.0
    for k, v in backend.TIMER.summary().items():
        ms = v["ms"] / v["calls"]
        tot += v["ms"] / reps
        print(f"   {k:40s} {ms:7.3f} ms  x{v['calls'] // reps}")
    print(f"   total per fwd+bwd: {tot:.3f} ms")
