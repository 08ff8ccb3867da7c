"""GPU: gradient parity of the fused SA MLP given IDENTICAL ReLU / arg-max decisions (float64 reference with the
decisions of the fused forward forced in) -- separates arithmetic error from decision flips."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ogc_b200 import segnet, sa_fused
import pointnet2.pointnet2 as ops


def run(N, M, Cf, widths, B=4, chain=False):
    torch.manual_seed(N + Cf)
    xyz = torch.randn(B, N, 3, device="cuda")
    new_xyz = xyz[:, :M].contiguous()
    feat_pm = torch.randn(B, N, Cf, device="cuda")
    mlp = segnet.SharedMLP([Cf + 3] + widths).cuda()
    with torch.no_grad():
        for n_, p_ in mlp.named_parameters():
            if "gn.weight" in n_: p_.copy_(torch.randn_like(p_) * 0.5 + 0.8)
            if "gn.bias" in n_: p_.copy_(torch.randn_like(p_) * 0.3)
    dist, idx = ops.knn(64, new_xyz, xyz)
    idx = ops.clip_neighbours_by_radius(dist, idx, 1.2)
    L = len(widths)
    layers = [(getattr(mlp, f"layer{i}").conv.weight, getattr(mlp, f"layer{i}").normlayer.gn.weight, getattr(mlp, f"layer{i}").normlayer.gn.bias) for i in range(L)]
    probe = torch.randn(B, widths[-1], M, device="cuda")
    f2 = feat_pm.clone().requires_grad_(True)
    sa_fused.USE_CHAIN = chain
    out = sa_fused.fused_sa_mlp(xyz, new_xyz, f2, idx, layers)
    saved = out.grad_fn.saved_tensors
    sel, ysel = saved[4], saved[5]
    ys, sss = saved[6:6 + L], saved[6 + L:6 + 2 * L]
    (out * probe).sum().backward()
    mine = {"dfeat": f2.grad.clone()}
    for i in range(L):
        for j, nm in enumerate(("W", "gamma", "beta")):
            mine[f"{nm}{i}"] = layers[i][j].grad.clone(); layers[i][j].grad = None
    # ---- float64 reference with forced decisions ----
    S = 64
    f64 = feat_pm.double().transpose(1, 2).contiguous().requires_grad_(True)          # (B,Cf,N)
    gi = idx.long()
    def group(t):   # (B,C,N) -> (B,C,M,S)
        Bc, C, _ = t.shape
        return torch.gather(t.unsqueeze(2).expand(Bc, C, M, t.shape[2]), 3, gi.unsqueeze(1).expand(Bc, C, M, S))
    a = torch.cat([group(xyz.double().transpose(1, 2).contiguous()) - new_xyz.double().transpose(1, 2).unsqueeze(-1), group(f64)], 1)
    params64 = [[t.detach().double().requires_grad_(True) for t in lay] for lay in layers]
    flips = []
    for l in range(L):
        W, gm, bt = params64[l]
        y = torch.einsum("oc,bcms->boms", W.reshape(W.shape[0], -1), a)
        Bc, C = y.shape[:2]
        yg = y.reshape(Bc, 4, -1)
        mu, var = yg.mean(2, keepdim=True), yg.var(2, unbiased=False, keepdim=True)
        z = ((yg - mu) / torch.sqrt(var + 1e-5)).reshape_as(y) * gm.view(1, -1, 1, 1) + bt.view(1, -1, 1, 1)
        m_ours = (sss[l][..., 0].double().view(Bc, C, 1) * ys[l].double() + sss[l][..., 1].double().view(Bc, C, 1) > 0).reshape_as(y)
        flips.append(int(((z > 0) != m_ours).sum()))
        a = z * m_ours
    selL = sel.long().clamp(max=S - 1).unsqueeze(-1)
    pooled = torch.gather(a, 3, selL).squeeze(-1) * (sel != 255)
    nat = a.max(dim=3)
    argflips = int(((nat.indices != sel.long()) & (sel != 255)).sum())
    (pooled * probe.double()).sum().backward()
    ref = {"dfeat": f64.grad.transpose(1, 2)}
    for i in range(L):
        for j, nm in enumerate(("W", "gamma", "beta")):
            ref[f"{nm}{i}"] = params64[i][j].grad.reshape(mine[f"{nm}{i}"].shape)
    print(f"== N={N} M={M} Cf={Cf} widths={widths} chain={chain}: ReLU decisions differing from fp64 per layer {flips}, arg-max differing {argflips}; "
          f"out err {float((out.double() - pooled).abs().max()):.2e}")
    for k in mine:
        d = (mine[k].double() - ref[k]).abs()
        print(f"   {k:7s} fro-rel {float(d.norm() / ref[k].norm()):.2e}   max-rel {float(d.max() / ref[k].abs().max()):.2e}")


run(2048, 1024, 96, [64, 64, 128])
run(1024, 512, 128, [128, 128, 256])
run(4096, 1024, 3, [32, 32, 64])
