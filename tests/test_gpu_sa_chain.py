"""GPU: the round-2 set-abstraction kernels against the round-1 per-layer kernels and against the torch-composed reference
expression (utils/pointnet2_util.py:33-44).
  * chained forward (csrc/sa_chain_fwd.cu: positions on the MMA's M axis, layers chained through tensor memory): pooled
    output, GroupNorm scale/shift, arg-max positions, stored pre-norm tensors;
  * tensor-map-TMA kernels: forward (sa_fwd_tma.cu), input gradient (sa_chain_bwd.cu positions-on-M, sa_dx_tma.cu
    channel-major), weight gradient (sa_dw_tma.cu) -- every gradient of the block, ONE forward and two backward passes
    over the same saved tensors (two forward runs can differ in a ReLU / arg-max decision)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

SHAPES = [
    (512, 128, 3, [32, 32, 32]),          # SA1 scale a (features = raw coordinates): "full" chain, small-row producer
    (512, 128, 3, [32, 32, 64]),          # SA1 scale b
    (400, 100, 96, [64, 64, 128]),        # SA2: "full" chain (3 passes, nothing stored)
    (300, 70, 128, [128, 128, 256]),      # SA3: one layer per launch (weights exceed one SM), last layer in 2 channel slices
    (256, 64, 16, [64, 64]),              # two-layer MLP, K1 = 16 (8-column operand tail)
    (256, 64, 40, [32, 96, 160]),         # odd widths: 40-channel rows (32 + 8 tail), group sizes 8 / 24 / 40
]


def _setup(N, M, Cf, widths, B=3):
    from ogc_b200 import segnet
    import pointnet2.pointnet2 as ops
    torch.manual_seed(N + Cf)
    xyz = torch.randn(B, N, 3, device="cuda")
    new_xyz = xyz[:, :M].contiguous()
    feat_pm = torch.randn(B, N, Cf, device="cuda")
    mlp = segnet.SharedMLP([Cf + 3] + widths).cuda()
    with torch.no_grad():
        for n_, p_ in mlp.named_parameters():
            if "gn.weight" in n_:
                p_.copy_(torch.randn_like(p_) * 0.5 + 0.8)     # some negative gammas exercise the min branch
            if "gn.bias" in n_:
                p_.copy_(torch.randn_like(p_) * 0.3)
    dist, idx = ops.knn(64, new_xyz, xyz)
    idx = ops.clip_neighbours_by_radius(dist, idx, 1.2)
    layers = [(getattr(mlp, f"layer{i}").conv.weight, getattr(mlp, f"layer{i}").normlayer.gn.weight,
               getattr(mlp, f"layer{i}").normlayer.gn.bias) for i in range(mlp.n_layers)]
    return xyz, new_xyz, feat_pm, idx, mlp, layers


@pytest.mark.parametrize("N,M,Cf,widths", SHAPES)
def test_chain_forward_matches_per_layer_kernels(b200, N, M, Cf, widths):
    from ogc_b200 import sa_fused
    xyz, new_xyz, feat_pm, idx, mlp, layers = _setup(N, M, Cf, widths)
    sa_fused.USE_CHAIN = True
    try:
        assert sa_fused._chain_plan(M, 64, Cf, widths) is not None
    finally:
        sa_fused.USE_CHAIN = False
    saved = {}
    for chain in (False, True):
        sa_fused.USE_CHAIN, sa_fused.STORE_Y = chain, True
        try:
            out = sa_fused.fused_sa_mlp(xyz, new_xyz, feat_pm.clone().requires_grad_(True), idx, layers)
        finally:
            sa_fused.USE_CHAIN = False
        saved[chain] = [out.detach().clone()] + [t.clone() for t in out.grad_fn.saved_tensors[4:]]
    assert len(saved[False]) == len(saved[True])
    for a, b in zip(saved[False], saved[True]):
        assert a.shape == b.shape
        if a.dtype == torch.uint8:
            assert float((a != b).float().mean()) < 1e-4       # arg-max positions (ties aside) identical
        else:
            assert float((a - b).abs().max()) <= 2e-5 * max(1.0, float(a.abs().max())), float((a - b).abs().max())


@pytest.mark.parametrize("N,M,Cf,widths", SHAPES)
def test_chain_matches_composed_reference_expression(b200, N, M, Cf, widths):
    """Values and every gradient against grouping_operation + SharedMLP + max in torch (the reference's op sequence)."""
    import pointnet2.pointnet2 as ops
    from ogc_b200.sa_fused import fused_sa_mlp
    xyz, new_xyz, feat_pm, idx, mlp, layers = _setup(N, M, Cf, widths)
    probe = torch.randn(3, widths[-1], M, device="cuda")
    f1 = feat_pm.transpose(1, 2).contiguous().requires_grad_(True)
    grouped = torch.cat([ops.grouping_operation(xyz.transpose(1, 2).contiguous(), idx) - new_xyz.transpose(1, 2).unsqueeze(-1),
                         ops.grouping_operation(f1, idx)], dim=1)
    ref = mlp(grouped).max(dim=3).values
    (ref * probe).sum().backward()
    ref_grads = {n: p.grad.clone() for n, p in mlp.named_parameters()}
    ref_df = f1.grad.clone()
    mlp.zero_grad()
    f2 = feat_pm.clone().requires_grad_(True)
    from ogc_b200 import sa_fused
    sa_fused.USE_CHAIN = True
    try:
        out = fused_sa_mlp(xyz, new_xyz, f2, idx, layers)
        (out * probe).sum().backward()
    finally:
        sa_fused.USE_CHAIN = False
    err = float((out.detach() - ref.detach()).abs().max() / ref.detach().abs().max())
    assert err < 2e-5, err

    def fro(a, b):
        return float((a - b).norm() / b.norm().clamp_min(1e-12))
    assert fro(f2.grad.transpose(1, 2), ref_df) < 2e-3
    for n, p in mlp.named_parameters():
        assert fro(p.grad, ref_grads[n]) < 2e-3, (n, fro(p.grad, ref_grads[n]))


@pytest.mark.parametrize("N,M,Cf,widths", [(2048, 1024, 96, [64, 64, 128]), (1024, 512, 128, [128, 128, 256]),
                                           (4096, 2048, 3, [32, 32, 64])])
def test_chain_forward_many_tiles_per_cta(b200, N, M, Cf, widths):
    """KITTI-SF sizes with 16 clouds: dozens of tiles per persistent CTA (pipeline wrap-around, barrier phases)."""
    from ogc_b200 import sa_fused
    xyz, new_xyz, feat_pm, idx, mlp, layers = _setup(N, M, Cf, widths, B=16)
    outs = {}
    for chain in (False, True):
        sa_fused.USE_CHAIN = chain
        try:
            outs[chain] = sa_fused.fused_sa_mlp(xyz, new_xyz, feat_pm.clone().requires_grad_(True), idx, layers).detach()
        finally:
            sa_fused.USE_CHAIN = False
    torch.cuda.synchronize()
    assert float((outs[True] - outs[False]).abs().max()) <= 2e-5 * max(1.0, float(outs[False].abs().max()))


DX_SHAPES = [
    (400, 100, 96, [64, 64, 128]),        # SA2: dense 128 -> 64 (synthesised dz), 64 -> 64, scatter 96
    (300, 70, 128, [128, 128, 256]),      # SA3: 256 -> 128 as two 64-row launches (resident weights), 128 -> 128, scatter 128
    (1024, 512, 128, [128, 128, 128]),    # many tiles per CTA with the shallow rings of the 128-row layers
    (256, 64, 32, [32, 96, 160]),         # group sizes 8 / 24 inside a 16-column piece; scatter 32
    (256, 64, 160, [64, 64]),             # scatter in two row blocks (128 + 32)
]


@pytest.mark.parametrize("N,M,Cf,widths", DX_SHAPES)
def test_chain_dx_matches_per_layer_kernels(b200, N, M, Cf, widths):
    """ogc_sa_chain_dx against ogc_sa_mlp_layer_dx_tc / ogc_sa_mlp_layer_dx: every gradient of the block."""
    from ogc_b200 import sa_fused
    xyz, new_xyz, feat_pm, idx, mlp, layers = _setup(N, M, Cf, widths)
    probe = torch.randn(3, widths[-1], M, device="cuda")
    # ONE forward, two backward passes over the same saved tensors: two forward runs can differ in a ReLU / arg-max
    # decision (GroupNorm statistics are accumulated atomically), which would be charged to the backward kernel
    f = feat_pm.clone().requires_grad_(True)
    out = sa_fused.fused_sa_mlp(xyz, new_xyz, f, idx, layers)
    loss = (out * probe).sum()
    wrt = [f] + list(mlp.parameters())
    res, default = {}, sa_fused.USE_CHAIN_DX
    for chain_dx in (False, True):
        sa_fused.USE_CHAIN_DX = chain_dx
        try:
            res[chain_dx] = [g.clone() for g in torch.autograd.grad(loss, wrt, retain_graph=True)]
        finally:
            sa_fused.USE_CHAIN_DX = default
    names = ["dfeat"] + [n for n, _ in mlp.named_parameters()]
    for name, a, b in zip(names, res[False], res[True]):
        rel = float((a - b).norm() / a.norm().clamp_min(1e-30))
        assert rel <= 2e-5, (name, rel)
        assert float((a - b).abs().max()) <= 1e-4 * float(a.abs().max()), (name, float((a - b).abs().max()))


DW_SHAPES = [
    (400, 100, 96, [64, 64, 128], False),       # SA2: 64 -> 128 (synthesised dz, one M block), 64 -> 64
    (300, 70, 128, [128, 128, 256], False),     # SA3: 128 -> 256 (two M blocks, two stages), 128 -> 128
    (1024, 512, 128, [128, 128, 128], False),   # many tiles per CTA
    (512, 128, 3, [32, 32, 64], True),          # SA1's narrow layers through the same kernel (padded M)
    (256, 64, 40, [32, 96, 160], False),        # 32 -> 96 dense, 96 -> 160 synthesised
]


@pytest.mark.parametrize("N,M,Cf,widths,narrow", DW_SHAPES)
def test_dw_tma_matches_per_layer_kernels(b200, N, M, Cf, widths, narrow):
    """ogc_sa_dw_tma against ogc_sa_mlp_layer_dw_tc / ogc_sa_mlp_narrow_dw: ONE forward, two backward passes.  The
    operands' hi parts are truncated (by the tensor core) instead of rounded: products carry 2^-20 instead of 2^-22."""
    from ogc_b200 import sa_fused
    xyz, new_xyz, feat_pm, idx, mlp, layers = _setup(N, M, Cf, widths)
    probe = torch.randn(3, widths[-1], M, device="cuda")
    f = feat_pm.clone().requires_grad_(True)
    out = sa_fused.fused_sa_mlp(xyz, new_xyz, f, idx, layers)
    loss = (out * probe).sum()
    wrt = [f] + list(mlp.parameters())
    res, default, default_n = {}, sa_fused.USE_DW_TMA, sa_fused.DW_TMA_NARROW
    for tma in (False, True):
        sa_fused.USE_DW_TMA, sa_fused.DW_TMA_NARROW = tma, narrow
        try:
            res[tma] = [g.clone() for g in torch.autograd.grad(loss, wrt, retain_graph=True)]
        finally:
            sa_fused.USE_DW_TMA, sa_fused.DW_TMA_NARROW = default, default_n
    names = ["dfeat"] + [n for n, _ in mlp.named_parameters()]
    for name, a, b in zip(names, res[False], res[True]):
        rel = float((a - b).norm() / a.norm().clamp_min(1e-30))
        assert rel <= 2e-5, (name, rel)
        assert float((a - b).abs().max()) <= 1e-4 * float(a.abs().max()), (name, float((a - b).abs().max()))


@pytest.mark.parametrize("N,M,Cf,widths,narrow", DW_SHAPES)
def test_fwd_tma_matches_per_layer_kernels(b200, N, M, Cf, widths, narrow):
    """ogc_sa_fwd_tma against ogc_sa_mlp_layer_fwd_tc / ogc_sa_mlp_narrow_fwd: pooled output, stored pre-norm tensors,
    GroupNorm scale / shift, arg-max bytes (ties / near-ties aside)."""
    from ogc_b200 import sa_fused
    xyz, new_xyz, feat_pm, idx, mlp, layers = _setup(N, M, Cf, widths)
    saved, default, default_n = {}, sa_fused.USE_FWD_TMA, sa_fused.FWD_TMA_NARROW
    for tma in (False, True):
        sa_fused.USE_FWD_TMA, sa_fused.FWD_TMA_NARROW = tma, narrow
        try:
            out = sa_fused.fused_sa_mlp(xyz, new_xyz, feat_pm.clone().requires_grad_(True), idx, layers)
        finally:
            sa_fused.USE_FWD_TMA, sa_fused.FWD_TMA_NARROW = default, default_n
        saved[tma] = [out.detach().clone()] + [t.clone() for t in out.grad_fn.saved_tensors[4:]]
    assert len(saved[False]) == len(saved[True])
    for i, (a, b) in enumerate(zip(saved[False], saved[True])):
        assert a.shape == b.shape
        if a.dtype == torch.uint8:
            assert float((a != b).float().mean()) < 1e-3, i
        else:
            assert float((a - b).abs().max()) <= 2e-5 * max(1.0, float(a.abs().max())), (i, float((a - b).abs().max()))


@pytest.mark.parametrize("mode", [True, "synth"])
@pytest.mark.parametrize("N,M,Cf,widths,narrow", DW_SHAPES)
def test_dx_tma_matches_per_layer_kernels(b200, N, M, Cf, widths, narrow, mode):
    """ogc_sa_dx_tma against the per-layer input-gradient kernels: ONE forward, two backward passes."""
    from ogc_b200 import sa_fused
    xyz, new_xyz, feat_pm, idx, mlp, layers = _setup(N, M, Cf, widths)
    probe = torch.randn(3, widths[-1], M, device="cuda")
    f = feat_pm.clone().requires_grad_(True)
    out = sa_fused.fused_sa_mlp(xyz, new_xyz, f, idx, layers)
    loss = (out * probe).sum()
    wrt = [f] + list(mlp.parameters())
    res, d_tma, d_chain, d_nw = {}, sa_fused.USE_DX_TMA, sa_fused.USE_CHAIN_DX, sa_fused.USE_NARROW
    for new in (False, True):
        sa_fused.USE_DX_TMA, sa_fused.USE_CHAIN_DX = (mode if new else False), False
        sa_fused.USE_NARROW = d_nw and not (new and narrow)      # route SA1's narrow layers through the new kernel as well
        try:
            res[new] = [g.clone() for g in torch.autograd.grad(loss, wrt, retain_graph=True)]
        finally:
            sa_fused.USE_DX_TMA, sa_fused.USE_CHAIN_DX, sa_fused.USE_NARROW = d_tma, d_chain, d_nw
    names = ["dfeat"] + [n for n, _ in mlp.named_parameters()]
    for name, a, b in zip(names, res[False], res[True]):
        rel = float((a - b).norm() / a.norm().clamp_min(1e-30))
        assert rel <= 2e-5, (name, rel)
        assert float((a - b).abs().max()) <= 1e-4 * float(a.abs().max()), (name, float((a - b).abs().max()))
