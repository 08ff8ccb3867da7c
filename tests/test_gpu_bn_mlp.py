"""Fused BatchNorm shared MLP of the FlowStep3D blocks (ogc_b200/bn_fused.py over csrc/bn_mlp.cu and the pointwise
contraction kernels) against the torch expression of the reference lines it replaces
(utils/flowstep3d_util.py:126-137 / :52-64: Conv2d 1x1 -> BatchNorm2d (training mode) -> ReLU, x L, max over nsample),
evaluated in fp64 on the same device: outputs, input gradient, weight / BatchNorm gradients, running estimates."""
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

# (B, Cin, M, S, widths, use_act): the block shapes of models/flownet_ogcdr.py:146-187 (scaled down in M)
SHAPES = [
    (3, 6, 96, 16, [32, 32, 32], True),        # encoder_loc.sa1
    (2, 35, 64, 16, [64, 64, 64], True),       # encoder_loc.sa2 (odd input width)
    (2, 67, 48, 16, [128, 128, 128], True),    # encoder_glob.sa1
    (2, 131, 40, 8, [64, 64, 64], True),       # local_corr_layer
    (3, 35, 50, 4, [16, 16, 16], True),        # flow_conv2 (16-wide layers, nsample 4, M not a multiple of 4)
    (2, 214, 64, 4, [64], False),              # gru.convz / convr / convq: bare convolution
    (1, 67, 33, 5, [64, 32], True),            # nsample not a multiple of 4: scalar paths
]


def _block(cin, widths, seed):
    g = torch.Generator().manual_seed(seed)
    convs, bns = nn.ModuleList(), nn.ModuleList()
    last = cin
    for c in widths:
        convs.append(nn.Conv2d(last, c, 1, bias=False))
        bn = nn.BatchNorm2d(c)
        with torch.no_grad():
            bn.weight.add_(0.3 * torch.randn(c, generator=g))      # some negative-leaning scales, non-zero shifts
            bn.bias.add_(0.2 * torch.randn(c, generator=g))
        bns.append(bn)
        last = c
    return convs, bns


def _reference(x, convs, bns, use_act):
    """fp64 torch: the reference's loop."""
    for conv, bn in zip(convs, bns):
        x = F.conv2d(x, conv.weight)
        if use_act:
            x = F.relu(F.batch_norm(x, bn.running_mean, bn.running_var, bn.weight, bn.bias, True, bn.momentum, bn.eps))
    return x.max(dim=-1).values


@pytest.mark.parametrize("tma", [True, "bwd", False])
@pytest.mark.parametrize("B,cin,M,S,widths,use_act", SHAPES)
def test_fused_bn_mlp_matches_fp64_torch(b200, monkeypatch, B, cin, M, S, widths, use_act, tma):
    """tma: dense inner layers through the TMA-staged tensor-core kernels (3xTF32) where the shape allows (positions a
    multiple of 128, widths multiples of 32) / their gradients only / everything through the fp32 SIMT kernels."""
    from ogc_b200 import bn_fused
    monkeypatch.setattr(bn_fused, "USE_TMA", tma)
    assert bn_fused.supported(widths, S)
    g = torch.Generator().manual_seed(B * 1000 + cin)
    x = (torch.randn(B, cin, M, S, generator=g) * 0.7 + 0.1).cuda()
    probe = torch.randn(B, widths[-1], M, generator=g).cuda()
    convs, bns = _block(cin, widths, seed=cin)
    convs, bns = convs.cuda(), bns.cuda()
    import copy
    convs64, bns64 = copy.deepcopy(convs).double(), copy.deepcopy(bns).double()

    x32 = x.clone().requires_grad_(True)
    out = bn_fused.fused_bn_mlp(x32, convs, bns if use_act else None)
    (out * probe).sum().backward()

    x64 = x.double().requires_grad_(True)
    ref = _reference(x64, convs64, bns64, use_act)
    (ref * probe.double()).sum().backward()

    def rel(a, b):
        return float((a.double() - b).norm() / b.norm().clamp_min(1e-30))

    e_out = float((out.double() - ref).abs().max() / ref.abs().max())
    print(f"out {e_out:.2e}")
    assert e_out <= 2e-5
    e_x = rel(x32.grad, x64.grad)
    print(f"dx {e_x:.2e}")
    assert e_x <= 2e-3
    for l, (c32, c64) in enumerate(zip(convs, convs64)):
        e = rel(c32.weight.grad, c64.weight.grad)
        print(f"dW{l} {e:.2e}")
        assert e <= 2e-3
    if use_act:
        for l, (b32, b64) in enumerate(zip(bns, bns64)):
            eg, eb = rel(b32.weight.grad, b64.weight.grad), rel(b32.bias.grad, b64.bias.grad)
            print(f"dgamma{l} {eg:.2e} dbeta{l} {eb:.2e}")
            assert eg <= 2e-3 and eb <= 2e-3
            assert rel(b32.running_mean, b64.running_mean) <= 1e-5 and rel(b32.running_var, b64.running_var) <= 1e-5
            assert int(b32.num_batches_tracked) == 1
    else:
        assert all(b.weight.grad is None for b in bns)


def test_fused_bn_mlp_skips_the_gradient_of_leading_rows_on_request(b200):
    """no_grad_rows = 3 (the centred coordinates of a leaf cloud): rows 3.. of the input gradient and every parameter
    gradient are what the full computation gives, rows 0..2 are zeros."""
    from ogc_b200 import bn_fused
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 67, 64, 16, generator=g).cuda()
    probe = torch.randn(2, 64, 64, generator=g).cuda()
    convs, bns = _block(67, [64, 64], seed=4)
    convs, bns = convs.cuda(), bns.cuda()
    params = list(convs.parameters()) + list(bns.parameters())
    res = []
    for skip in (0, 3):
        xi = x.clone().requires_grad_(True)
        for p in params:
            p.grad = None
        (bn_fused.fused_bn_mlp(xi, convs, bns, skip) * probe).sum().backward()
        res.append((xi.grad.clone(), [p.grad.clone() for p in params]))
    (gx0, gp0), (gx3, gp3) = res
    assert float(gx3[:, :3].abs().max()) == 0.0
    assert float((gx3[:, 3:] - gx0[:, 3:]).norm() / gx0[:, 3:].norm()) <= 1e-5
    for a, b in zip(gp0, gp3):
        assert float((a - b).norm() / a.norm().clamp_min(1e-30)) <= 1e-4


def test_fused_bn_mlp_accumulates_into_existing_grads(b200):
    """The trainer's backward: gradients are added straight into the parameters' .grad (sa_fused.grad_targets); a block
    used twice in one forward (the reference re-applies its encoders every GRU iteration) sums both uses."""
    from ogc_b200 import bn_fused
    g = torch.Generator().manual_seed(3)
    xa, xb = (torch.randn(2, 35, 64, 8, generator=g).cuda() for _ in range(2))
    convs, bns = _block(35, [32, 32], seed=9)
    convs, bns = convs.cuda(), bns.cuda()
    params = [p for p in list(convs.parameters()) + list(bns.parameters())]
    (bn_fused.fused_bn_mlp(xa, convs, bns).sum() + bn_fused.fused_bn_mlp(xb, convs, bns).square().sum()).backward()
    want = [p.grad.clone() for p in params]
    for p in params:
        p.grad = torch.zeros_like(p)
    (bn_fused.fused_bn_mlp(xa, convs, bns).sum() + bn_fused.fused_bn_mlp(xb, convs, bns).square().sum()).backward()
    for p, w in zip(params, want):
        assert float((p.grad - w).norm() / w.norm().clamp_min(1e-30)) <= 1e-4


def test_flownet_fused_blocks_match_composed_blocks(b200):
    """Whole network, one forward + backward in training mode: fused BatchNorm blocks vs torch matmul + BatchNorm2d."""
    from ogc_b200 import flownet
    from tests.golden.cases import CASES, build_my_flownet, make_inputs
    case = CASES["flownet_ogcdr_512"]
    inp = {k: v.cuda() for k, v in make_inputs(case).items()}

    def run(fused):
        prev, flownet.USE_FUSED_MLP = flownet.USE_FUSED_MLP, fused
        try:
            net = build_my_flownet(case).cuda()
            preds = net(inp["pc1"], inp["pc2"], inp["pc1"], inp["pc2"], iters=2)
            (preds[0].square().sum() + preds[1].square().sum()).backward()
            grads = {n: p.grad.clone() for n, p in net.named_parameters() if p.grad is not None}
            stats = {n: b.clone() for n, b in net.named_buffers() if "running" in n}
            return [p.detach() for p in preds], grads, stats
        finally:
            flownet.USE_FUSED_MLP = prev

    pf, gf, sf = run(True)
    pc, gc, sc = run(False)
    for a, b in zip(pf, pc):       # untrained network: flows of magnitude ~2; 1e-4 of the range
        err = float((a - b).abs().max() / b.abs().max())
        print(f"flow prediction: {err:.2e} of the range")
        assert err <= 1e-4
    assert set(gf) == set(gc)
    worst = max(float((gf[n] - gc[n]).norm() / gc[n].norm().clamp_min(1e-12)) for n in gc)
    print(f"worst relative gradient difference {worst:.2e}")
    assert worst <= 5e-3
    for n in sc:
        assert float((sf[n] - sc[n]).abs().max()) <= 1e-4 * max(1.0, float(sc[n].abs().max())), n
