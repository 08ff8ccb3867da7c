"""CPU tests of the drop-in operator layer (pointnet2/pointnet2.py) running on the oracle back-end:
public names, return conventions, autograd, and composite modules against plain-torch definitions.
Mirrors what the reference's star-importers rely on."""
import numpy as np
import pytest
import torch


def test_public_names_match_reference_module():
    import pointnet2.pointnet2 as ops
    for name in ["torch", "nn", "Function", "Variable", "Tuple", "gather_nd",
                 "FurthestPointSampling", "furthest_point_sample", "GatherOperation", "gather_operation",
                 "KNN", "knn", "ThreeNN", "three_nn", "ThreeInterpolate", "three_interpolate",
                 "GroupingOperation", "grouping_operation", "BallQuery", "ball_query",
                 "QueryAndGroup", "GroupAll"]:
        assert hasattr(ops, name), name


def test_default_backend_is_the_cuda_library_and_rejects_cpu():
    """Without a test installing the oracle, the API goes to libogc_b200 and refuses CPU tensors."""
    from ogc_b200 import backend
    import pointnet2.pointnet2 as ops
    prev = backend.set_backend(None)
    try:
        with pytest.raises(RuntimeError, match="CUDA tensor"):
            ops.furthest_point_sample(torch.zeros(1, 4, 3), 2)
        assert backend.get_backend().name == "b200"
    finally:
        backend.set_backend(prev)


def test_gather_nd(oracle_ops):
    p = torch.arange(2 * 5 * 3, dtype=torch.float32).view(2, 5, 3)
    idx = torch.tensor([[0, 4], [3, 3]])
    out = oracle_ops.gather_nd(p, idx)
    assert out.shape == (2, 2, 3) and torch.equal(out[1, 0], p[1, 3])
    out_t = oracle_ops.gather_nd(p.transpose(1, 2).contiguous(), idx, t=True)
    assert torch.equal(out_t, out.transpose(1, 2))


def test_knn_returns_sqrt_and_fresh_writable_tensors(oracle_ops, oracle):
    rng = np.random.default_rng(0)
    a = torch.from_numpy(rng.normal(size=(2, 30, 3)).astype(np.float32))
    b = torch.from_numpy(rng.normal(size=(2, 50, 3)).astype(np.float32))
    dist, idx = oracle_ops.knn(4, a, b)
    d2, idx_ref = oracle.knn(4, a, b)
    assert idx.dtype == torch.int32 and torch.equal(idx, idx_ref)
    assert torch.equal(dist, torch.sqrt(d2))
    idx[dist > 0.5] = 0          # callers mutate in place (reference pointnet2.py:286, seg_loss_unsup.py:122)
    assert not dist.requires_grad and not idx.requires_grad


def test_grouping_and_interpolation_gradients_match_torch(oracle_ops):
    rng = np.random.default_rng(1)
    B, C, N, M, S = 2, 4, 25, 6, 5
    f = torch.from_numpy(rng.normal(size=(B, C, N)).astype(np.float32)).requires_grad_(True)
    idx = torch.from_numpy(rng.integers(0, N, size=(B, M, S)).astype(np.int64))   # any int dtype is accepted
    out = oracle_ops.grouping_operation(f, idx)
    w = torch.from_numpy(rng.normal(size=out.shape).astype(np.float32))
    (out * w).sum().backward()
    f2 = f.detach().clone().requires_grad_(True)
    ref = torch.gather(f2.unsqueeze(2).expand(B, C, M, N), 3, idx.unsqueeze(1).expand(B, C, M, S))
    (ref * w).sum().backward()
    assert torch.equal(out, ref)
    torch.testing.assert_close(f.grad, f2.grad, rtol=1e-5, atol=1e-6)

    n = 8
    idx3 = torch.from_numpy(rng.integers(0, N, size=(B, n, 3)).astype(np.int32))
    w3 = torch.from_numpy(rng.random(size=(B, n, 3)).astype(np.float32))
    f.grad = None
    o = oracle_ops.three_interpolate(f, idx3, w3)
    (o ** 2).sum().backward()
    f2.grad = None
    picked = torch.gather(f2.unsqueeze(2).expand(B, C, n, N), 3, idx3.long().unsqueeze(1).expand(B, C, n, 3))
    o2 = (picked * w3.unsqueeze(1)).sum(-1)
    (o2 ** 2).sum().backward()
    torch.testing.assert_close(o, o2, rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(f.grad, f2.grad, rtol=1e-5, atol=1e-5)

    f.grad = None
    g = oracle_ops.gather_operation(f, idx[:, :, 0].int().contiguous())
    g.sum().backward()
    cnt = torch.zeros(B, N).scatter_add_(1, idx[:, :, 0], torch.ones(B, M))
    torch.testing.assert_close(f.grad, cnt.unsqueeze(1).expand(B, C, N))


def test_query_and_group_definition(oracle_ops):
    """QueryAndGroup = kNN + radius clip on sqrt distance + group + centre + concat (reference :273-301)."""
    rng = np.random.default_rng(2)
    B, N, M, S, C = 2, 80, 10, 8, 3
    xyz = torch.from_numpy(rng.normal(size=(B, N, 3)).astype(np.float32))
    new_xyz = xyz[:, :M].contiguous()
    feats = torch.from_numpy(rng.normal(size=(B, C, N)).astype(np.float32))
    radius = 0.9
    out, gx = oracle_ops.QueryAndGroup(radius, S)(xyz, new_xyz, feats)
    assert out.shape == (B, 3 + C, M, S) and gx.shape == (B, 3, M, S)
    d = torch.cdist(new_xyz.double(), xyz.double())
    order = torch.argsort(d, dim=2, stable=True)[:, :, :S]
    dsel = torch.gather(d, 2, order)
    order = torch.where(dsel > radius, order[:, :, :1], order)
    ref_xyz = torch.gather(xyz.unsqueeze(1).expand(B, M, N, 3), 2, order.unsqueeze(-1).expand(B, M, S, 3))
    ref_xyz = (ref_xyz - new_xyz.unsqueeze(2)).permute(0, 3, 1, 2)
    ref_f = torch.gather(feats.unsqueeze(2).expand(B, C, M, N), 3, order.unsqueeze(1).expand(B, C, M, S))
    torch.testing.assert_close(out[:, :3], ref_xyz, rtol=0, atol=0)
    torch.testing.assert_close(out[:, 3:], ref_f, rtol=0, atol=0)
    assert torch.equal(gx, out[:, :3])
    # radius None -> plain kNN grouping; no features -> xyz only
    out2, _ = oracle_ops.QueryAndGroup(None, S)(xyz, new_xyz)
    assert out2.shape == (B, 3, M, S)
    ga, gxa = oracle_ops.GroupAll()(xyz, None, feats)
    assert ga.shape == (B, 3 + C, 1, N) and gxa.shape == (B, 3, 1, N)


def test_ball_query_and_fps_through_api(oracle_ops, oracle):
    rng = np.random.default_rng(3)
    xyz = torch.from_numpy(rng.normal(size=(2, 200, 3)).astype(np.float32))
    sel = oracle_ops.furthest_point_sample(xyz, 50)
    assert sel.dtype == torch.int32 and sel.shape == (2, 50) and (sel[:, 0] == 0).all()
    new_xyz = oracle_ops.gather_nd(xyz, sel.long()).contiguous()
    idx = oracle_ops.ball_query(0.7, 12, xyz, new_xyz)
    assert torch.equal(idx, oracle.ball_query(0.7, 12, xyz, new_xyz))
    assert (idx[:, :, 0] <= sel).all()   # the first hit is the LOWEST index in the ball, at most the centre's own


def test_flow_block_fusion_is_gpu_only_host_logic(monkeypatch):
    """Host logic of the FlowStep3D block fusion (ogc_b200/bn_fused.py, flownet._fused_mlp_ok): CPU tensors, eval-mode
    BatchNorm and unsupported widths take the torch-composed path (there is no CPU fallback INSIDE the fused path: it
    would raise on the missing device pointers); the tensor-core switch parses its three settings."""
    import importlib
    import torch
    import torch.nn as nn
    from ogc_b200 import bn_fused, flownet
    assert bn_fused.supported([16, 32, 64, 128], 16) and bn_fused.supported([256], 4)
    assert not bn_fused.supported([24], 16) and not bn_fused.supported([512], 16) and not bn_fused.supported([64], 255)
    sa = flownet.FlowSA(8, 4, 3, [32, 32])
    x = torch.zeros(2, 6, 8, 4)
    assert not flownet._fused_mlp_ok(x, sa.mlp_convs, sa.mlp_bns, True)            # CPU tensor
    for value, want in (("0", False), ("1", True), ("bwd", "bwd"), ("", "bwd")):
        monkeypatch.setenv("OGC_BN_TMA", value)
        assert importlib.reload(bn_fused).USE_TMA == want
    monkeypatch.delenv("OGC_BN_TMA")
    assert importlib.reload(bn_fused).USE_TMA == "bwd"
    assert isinstance(sa.mlp_bns[0], nn.BatchNorm2d)
