"""GPU: the failing integrated case (one forward, backward with per-layer vs chained dX kernels), all tensors reported."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from ogc_b200 import sa_fused
import test_gpu_sa_chain as t
N, M, Cf, widths = 1024, 512, 128, [128, 128, 128]
xyz, new_xyz, feat_pm, idx, mlp, layers = t._setup(N, M, Cf, widths)
probe = torch.randn(3, widths[-1], M, device="cuda")
f = feat_pm.clone().requires_grad_(True)
out = sa_fused.fused_sa_mlp(xyz, new_xyz, f, idx, layers)
loss = (out * probe).sum()
wrt = [f] + list(mlp.parameters())
names = ["dfeat"] + [n for n, _ in mlp.named_parameters()]
mode = os.environ.get("MODE", "1")
sa_fused.USE_CHAIN_DX = False
sa_fused.DEBUG_KEEP = []
ref = [g.clone() for g in torch.autograd.grad(loss, wrt, retain_graph=True)]
keep_ref = sa_fused.DEBUG_KEEP
sa_fused.DEBUG_KEEP = None
ref2 = [g.clone() for g in torch.autograd.grad(loss, wrt, retain_graph=True)]
print("old vs old:", {n: f"{float((a - b).norm() / a.norm()):.1e}" for n, a, b in zip(names, ref, ref2)})
sa_fused.USE_CHAIN_DX = {"1": True}.get(mode, mode)
for r in range(int(sys.argv[1]) if len(sys.argv) > 1 else 10):
    sa_fused.DEBUG_KEEP = []
    got = [g.clone() for g in torch.autograd.grad(loss, wrt, retain_graph=True)]
    for (l, dzr, abr, cfr), (l2, dzg, abg, cfg) in zip(keep_ref, sa_fused.DEBUG_KEEP):
        d = (dzr - dzg).abs()
        nz = (d > 1e-3 * dzr.abs().max()).nonzero()
        if len(nz):
            tiles = sorted(set((int(a), int(c) // 128) for a, _, c in nz.tolist()))
            print(f"   run {r} layer {l}: dz_prev bad {len(nz)} coef diff {float((cfr - cfg).abs().max()):.1e} ab diff {float((abr - abg).abs().max()):.1e} (sample, tile): {tiles[:12]} "
                  f"chan {int(nz[:, 1].min())}-{int(nz[:, 1].max())} pos in tile {int((nz[:, 2] % 128).min())}-{int((nz[:, 2] % 128).max())}")
            b0, t0 = tiles[0]
            blk = d[b0, :, t0 * 128:(t0 + 1) * 128] > 1e-3 * dzr.abs().max()
            print("      bad per 16-channel piece:", blk.view(-1, 16, 128).any(-1).sum(-1).tolist() if False else blk.view(8, 16, 128).flatten(1).sum(1).tolist(),
                  " bad per 32-position quadrant:", blk.view(128, 4, 32).permute(1, 0, 2).flatten(1).sum(1).tolist())
    rels = {n: float((a - b).norm() / a.norm()) for n, a, b in zip(names, ref, got)}
    if max(rels.values()) > 1e-5:
        print("run", r, {n: f"{x:.1e}" for n, x in rels.items() if x > 1e-5})
        d = (ref[0] - got[0]).abs()
        nz = (d > 1e-3 * ref[0].abs().max()).nonzero()
        if len(nz):
            print("   dfeat bad", len(nz), "samples", sorted(set(nz[:, 0].tolist())), "points", int(nz[:, 1].min()), int(nz[:, 1].max()), "n distinct points", len(set(nz[:, 1].tolist())),
                  "chan", int(nz[:, 2].min()), int(nz[:, 2].max()))
    else:
        print("run", r, "ok", f"{max(rels.values()):.1e}")
