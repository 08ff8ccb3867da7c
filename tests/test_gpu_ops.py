"""GPU parity tests: libogc_b200.so (through the C ABI, via ogc_b200.backend.B200Backend) against the
CPU oracle on the same seeded inputs.  Bar: bit-exact for every integer output (FPS / KNN / three_nn /
ball_query indices), bit-exact for forward fp32 outputs (same rounding order), 1e-5 for the
atomically-accumulated gradients (summation order is unspecified in the reference too).
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def clouds(seed, b, n, kind="normal"):
    rng = np.random.default_rng(seed)
    if kind == "int":      # exact arithmetic, ties everywhere
        a = rng.integers(-3, 4, size=(b, n, 3)).astype(np.float32)
    elif kind == "scene":  # KITTI-like extents (tens of metres)
        a = (rng.random(size=(b, n, 3)) * np.array([50, 4, 30]) - np.array([25, 2, -5])).astype(np.float32)
    else:
        a = rng.normal(size=(b, n, 3)).astype(np.float32)
    return torch.from_numpy(a)


# ------------------------------------------------------------------------------------ FPS (K1)
@pytest.mark.parametrize("n,m,kind", [
    (1, 1, "normal"), (2, 2, "normal"), (3, 3, "int"), (5, 5, "int"), (17, 9, "normal"), (31, 31, "int"),
    (32, 16, "normal"), (33, 33, "int"), (100, 64, "int"), (512, 128, "normal"), (1000, 257, "normal"),
    (1024, 512, "scene"), (1500, 700, "int"), (2048, 1024, "scene"), (3000, 100, "normal"),
    (4096, 1024, "scene"), (5000, 333, "int"), (8192, 2048, "scene"), (10000, 200, "normal"),
    (16384, 128, "scene"), (20000, 64, "normal"),
    # raw-scene sizes: thread-block-cluster kernel (8 CTAs x 4 / 8 points per thread, 16 CTAs), then the single-CTA fallback
    (16385, 40, "int"), (30000, 150, "scene"), (50000, 300, "int"), (65536, 64, "normal"), (100000, 200, "scene"),
    (131072, 32, "normal"), (140000, 24, "normal"),
])
def test_fps_bit_exact(b200, oracle, n, m, kind):
    xyz = clouds(n + m, 3, n, kind)
    ref = oracle.fps(xyz, m)
    got = b200.fps(xyz.cuda(), m).cpu()
    assert got.dtype == torch.int32
    assert torch.equal(got, ref), f"first mismatch at {(got != ref).nonzero()[:3].tolist()}"


def test_fps_full_size_properties(b200):
    """KITTI-SF sizes (16 x 8192 -> 2048): properties that need no oracle."""
    xyz = clouds(1, 16, 8192, "scene").cuda()
    idx = b200.fps(xyz, 2048)
    assert (idx[:, 0] == 0).all()
    srt = idx.sort(dim=1).values
    assert (srt[:, 1:] != srt[:, :-1]).all(), "FPS picked a point twice"
    # greedy property: min distance of each new sample to the previous ones is non-increasing
    sel = torch.gather(xyz, 1, idx.long().unsqueeze(-1).expand(-1, -1, 3))
    d = torch.cdist(sel[:, :512].double(), sel[:, :512].double())
    tri = torch.tril(torch.ones(512, 512, device="cuda", dtype=torch.bool), -1)
    mins = torch.where(tri, d, torch.full_like(d, 1e9)).min(dim=2).values[:, 1:]
    assert (mins[:, 1:] <= mins[:, :-1] + 1e-9).all()
    # FPS of an FPS-ordered prefix is the identity
    again = b200.fps(sel.contiguous(), 1024)
    assert torch.equal(again.cpu(), torch.arange(1024, dtype=torch.int32).expand(16, -1))


# --------------------------------------------------------------------------- KNN / three_nn
@pytest.mark.parametrize("n,m,k,kind", [
    (1, 1, 1, "normal"), (7, 5, 8, "normal"), (40, 33, 3, "int"), (100, 257, 1, "normal"), (64, 1000, 16, "int"),
    (300, 2048, 32, "scene"), (130, 4096, 64, "scene"), (50, 3000, 64, "int"), (33, 9000, 32, "normal"),
    (20, 700, 100, "int"), (10, 17000, 64, "scene"), (9, 600, 200, "int"), (12, 5, 64, "normal"),
])
def test_knn_bit_exact(b200, oracle, n, m, k, kind):
    q, r = clouds(n * 3 + k, 2, n, kind), clouds(m * 5 + k, 2, m, kind)
    d2_ref, idx_ref = oracle.knn(k, q, r)
    d2, idx = b200.knn(k, q.cuda(), r.cuda())
    assert torch.equal(idx.cpu(), idx_ref)
    assert torch.equal(d2.cpu(), d2_ref)
    dist, idx2 = b200.knn(k, q.cuda(), r.cuda(), sqrt=True)
    assert torch.equal(idx2.cpu(), idx_ref)
    assert torch.equal(dist, torch.sqrt(d2))      # = what the reference does on the GPU (pointnet2.py:103)


def test_knn_unaligned_cloud_pointers(b200, oracle):
    """m*12 bytes not a multiple of 16 -> batches start at unaligned addresses (TMA head/tail path)."""
    for m in (5, 6, 7, 1001, 1002, 1003):
        q, r = clouds(m, 3, 20), clouds(m + 1, 3, m)
        _, idx_ref = oracle.knn(4, q, r)
        _, idx = b200.knn(4, q.cuda(), r.cuda())
        assert torch.equal(idx.cpu(), idx_ref), m


def test_knn_self_query_full_size(b200, oracle):
    """Smooth-loss shape (8192 x 8192, k=32): first 256 queries against the oracle + properties."""
    pc = clouds(2, 2, 8192, "scene")
    d2, idx = b200.knn(32, pc.cuda(), pc.cuda())
    d2_ref, idx_ref = oracle.knn(32, pc[:, :256].contiguous(), pc)
    assert torch.equal(idx[:, :256].cpu(), idx_ref)
    assert torch.equal(d2[:, :256].cpu(), d2_ref)
    assert (idx[:, :, 0].cpu() == torch.arange(8192)).all()           # nearest neighbour of a point is itself
    assert (d2[:, :, 1:] >= d2[:, :, :-1]).all()                      # ascending
    got = torch.gather(pc.cuda().unsqueeze(1).expand(-1, 8192, -1, -1), 2,
                       idx.long().unsqueeze(-1).expand(-1, -1, -1, 3))
    chk = ((got - pc.cuda().unsqueeze(2)) ** 2).sum(-1)
    torch.testing.assert_close(chk, d2, rtol=1e-4, atol=1e-5)          # distances belong to the indices


def test_knn_ignores_nan_inf(b200, oracle):
    q = torch.zeros(1, 2, 3)
    r = torch.tensor([[[1., 0, 0], [float("nan"), 0, 0], [float("inf"), 0, 0], [2., 0, 0]]])
    d2_ref, idx_ref = oracle.knn(4, q, r)
    d2, idx = b200.knn(4, q.cuda(), r.cuda())
    assert torch.equal(idx.cpu(), idx_ref)
    assert torch.equal(d2.cpu(), d2_ref)


@pytest.mark.parametrize("n,m,kind", [(5, 2, "normal"), (33, 1, "int"), (1024, 512, "scene"), (2048, 1024, "int"),
                                      (8192, 2048, "scene")])
def test_three_nn_bit_exact(b200, oracle, n, m, kind):
    q, r = clouds(n, 2, n, kind), clouds(m + 1, 2, m, kind)
    d2_ref, idx_ref = oracle.three_nn(q, r)
    d2, idx = b200.three_nn(q.cuda(), r.cuda())
    assert torch.equal(idx.cpu(), idx_ref)
    assert torch.equal(d2.cpu(), d2_ref)


# ------------------------------------------------------------------------------- ball query
@pytest.mark.parametrize("n,m,r,ns,kind", [
    (1, 1, 1.0, 4, "normal"), (50, 20, 0.5, 8, "normal"), (100, 100, 2.0, 64, "int"), (1000, 300, 1.5, 16, "int"),
    (4096, 512, 2.0, 64, "scene"), (8192, 700, 2.0, 64, "scene"), (9001, 100, 0.1, 8, "normal"),
    (300, 40, 100.0, 40, "normal"), (64, 10, 1e-6, 5, "normal"), (17000, 50, 3.0, 33, "scene"),
])
def test_ball_query_bit_exact(b200, oracle, n, m, r, ns, kind):
    xyz, c = clouds(n + 1, 2, n, kind), clouds(m + 2, 2, m, kind)
    ref = oracle.ball_query(r, ns, xyz, c)
    got = b200.ball_query(r, ns, xyz.cuda(), c.cuda())
    assert got.dtype == torch.int32
    assert torch.equal(got.cpu(), ref)


def test_ball_query_self_full_size(b200, oracle):
    pc = clouds(4, 2, 8192, "scene")
    got = b200.ball_query(2.0, 64, pc.cuda(), pc.cuda()).cpu()
    ref = oracle.ball_query(2.0, 64, pc, pc[:, :300].contiguous())
    assert torch.equal(got[:, :300], ref)
    # every returned neighbour is inside the ball and rows are "ascending then repeated-first"
    g = torch.gather(pc.unsqueeze(1).expand(-1, 8192, -1, -1), 2, got.long().unsqueeze(-1).expand(-1, -1, -1, 3))
    assert (((g - pc.unsqueeze(2)) ** 2).sum(-1) < 4.0 + 1e-4).all()


# ---------------------------------------------------------------- gather / group / interpolate
@pytest.mark.parametrize("B,C,N,M,S", [(2, 3, 100, 17, 5), (2, 10, 8192, 512, 32), (1, 96, 2048, 1024, 64),
                                       (3, 7, 33, 9, 3), (2, 1, 64, 64, 1)])
def test_group_points_forward_backward(b200, oracle, B, C, N, M, S):
    rng = np.random.default_rng(B + C + N)
    f = torch.from_numpy(rng.normal(size=(B, C, N)).astype(np.float32))
    idx = torch.from_numpy(rng.integers(0, N, size=(B, M, S)).astype(np.int32))
    assert torch.equal(b200.group_points(f.cuda(), idx.cuda()).cpu(), oracle.group_points(f, idx))
    go = torch.from_numpy(rng.normal(size=(B, C, M, S)).astype(np.float32))
    g = b200.group_points_grad(go.cuda(), idx.cuda(), N).cpu()
    torch.testing.assert_close(g, oracle.group_points_grad(go, idx, N), rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("B,C,N,M", [(2, 3, 100, 17), (2, 64, 4096, 1024), (1, 5, 7, 7), (2, 9, 1000, 1000)])
def test_gather_points_forward_backward(b200, oracle, B, C, N, M):
    rng = np.random.default_rng(C + N + M)
    f = torch.from_numpy(rng.normal(size=(B, C, N)).astype(np.float32))
    idx = torch.from_numpy(rng.integers(0, N, size=(B, M)).astype(np.int32))
    assert torch.equal(b200.gather_points(f.cuda(), idx.cuda()).cpu(), oracle.gather_points(f, idx))
    go = torch.from_numpy(rng.normal(size=(B, C, M)).astype(np.float32))
    g = b200.gather_points_grad(go.cuda(), idx.cuda(), N).cpu()
    torch.testing.assert_close(g, oracle.gather_points_grad(go, idx, N), rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("B,C,m,n", [(2, 3, 50, 20), (2, 256, 512, 1024), (1, 64, 2048, 8192), (3, 5, 3, 11)])
def test_three_interpolate_forward_backward(b200, oracle, B, C, m, n):
    rng = np.random.default_rng(C + m + n)
    f = torch.from_numpy(rng.normal(size=(B, C, m)).astype(np.float32))
    idx = torch.from_numpy(rng.integers(0, m, size=(B, n, 3)).astype(np.int32))
    w = torch.from_numpy(rng.random(size=(B, n, 3)).astype(np.float32))
    out = b200.three_interpolate(f.cuda(), idx.cuda(), w.cuda()).cpu()
    assert torch.equal(out, oracle.three_interpolate(f, idx, w))       # same fma order -> bit-exact
    go = torch.from_numpy(rng.normal(size=(B, C, n)).astype(np.float32))
    g = b200.three_interpolate_grad(go.cuda(), idx.cuda(), w.cuda(), m).cpu()
    torch.testing.assert_close(g, oracle.three_interpolate_grad(go, idx, w, m), rtol=1e-5, atol=1e-5)


# ------------------------------------------------------------------- operator layer on the GPU
def test_operator_layer_autograd_on_gpu(oracle):
    """pointnet2.pointnet2 (the drop-in API) on CUDA tensors: values + gradients vs. the same layer on
    the CPU oracle."""
    from ogc_b200 import backend
    import pointnet2.pointnet2 as ops
    rng = np.random.default_rng(0)
    xyz = torch.from_numpy(rng.normal(size=(2, 600, 3)).astype(np.float32))
    feats = torch.from_numpy(rng.normal(size=(2, 8, 600)).astype(np.float32))

    def run(dev):
        x = xyz.to(dev)
        f = feats.to(dev).detach().clone().requires_grad_(True)
        inds = ops.furthest_point_sample(x, 128)
        new_xyz = ops.gather_nd(x, inds.long())
        grouped, _ = ops.QueryAndGroup(0.8, 16)(x, new_xyz.contiguous(), f)
        pooled = grouped.max(dim=3).values                        # (B, 3+C, 128)
        dist, idx = ops.three_nn(x, new_xyz.contiguous())
        w = 1.0 / (dist + 1e-8)
        w = w / w.sum(dim=2, keepdim=True)
        up = ops.three_interpolate(pooled.contiguous(), idx, w.contiguous())
        bq = ops.ball_query(0.5, 8, x, new_xyz.contiguous())
        gathered = ops.gather_operation(f, inds)
        loss = (up ** 2).sum() + gathered.sum() * 0.1
        loss.backward()
        return inds.cpu(), grouped.detach().cpu(), up.detach().cpu(), bq.cpu(), f.grad.cpu()

    prev = backend.set_backend(oracle)
    try:
        ref = run("cpu")
    finally:
        backend.set_backend(prev)
    got = run("cuda")
    assert torch.equal(got[0], ref[0]) and torch.equal(got[3], ref[3])
    assert torch.equal(got[1], ref[1])
    torch.testing.assert_close(got[2], ref[2], rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(got[4], ref[4], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("n,k,r,kind", [(2048, 32, 1.0, "scene"), (4096, 64, 2.0, "scene"), (1500, 16, 0.3, "normal"), (700, 8, 1e-3, "int")])
def test_knn_bounded_equals_exact_knn_after_radius_clip(b200, n, k, r, kind):
    """ogc_knn_bounded: identical to the exact k-NN once neighbours beyond the radius are replaced by the nearest
    one (what QueryAndGroup / KnnLoss do with the result)."""
    import pointnet2.pointnet2 as ops
    pc = clouds(n + k, 2, n, kind).cuda()
    dist, idx = b200.knn(k, pc, pc, sqrt=True)
    bd, bidx = b200.knn_bounded(k, pc, pc, r)
    assert torch.equal(ops.clip_neighbours_by_radius(dist, idx, r), ops.clip_neighbours_by_radius(bd, bidx, r))
    inside = dist <= r
    assert torch.equal(bd[inside], dist[inside]) and (torch.isinf(bd) | (bd == dist))[~inside].all()
