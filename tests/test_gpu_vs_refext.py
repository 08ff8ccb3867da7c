"""GPU differential tests against the REFERENCE's own CUDA extension (oracle/_ref, built from
/root/reference by oracle/ref_build.py).  This is what pins the oracle: the reference has no golden
vectors, so (1) the CPU restatement and (2) the sm_100a kernels are both compared with the real
extension running on the same B200.  Skipped when the extension was not built."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def refext():
    from oracle import refext as R
    if not R.available():
        pytest.skip("oracle/_ref/pointnet2_cuda_ref.so not built (python oracle/ref_build.py)")
    return R.RefExtBackend()


def clouds(seed, b, n, kind):
    rng = np.random.default_rng(seed)
    if kind == "int":
        a = rng.integers(-3, 4, size=(b, n, 3)).astype(np.float32)
    elif kind == "scene":
        a = (rng.random(size=(b, n, 3)) * np.array([50, 4, 30]) - np.array([25, 2, -5])).astype(np.float32)
    else:
        a = rng.normal(size=(b, n, 3)).astype(np.float32)
    return torch.from_numpy(a)


@pytest.mark.parametrize("n,m,kind", [(3, 3, "int"), (33, 33, "int"), (100, 64, "int"), (1000, 300, "int"),
                                      (1500, 700, "normal"), (4096, 1024, "scene"), (5000, 400, "int"),
                                      (8192, 2048, "scene"), (8192, 512, "int"), (20000, 64, "normal")])
def test_fps_three_way(b200, oracle, refext, n, m, kind):
    xyz = clouds(n + m, 3, n, kind)
    ref = refext.fps(xyz.cuda(), m).cpu()
    assert torch.equal(oracle.fps(xyz, m), ref), "CPU restatement differs from the reference extension"
    assert torch.equal(b200.fps(xyz.cuda(), m).cpu(), ref), "sm_100a kernel differs from the reference extension"


@pytest.mark.parametrize("n,m,k,kind", [(40, 33, 3, "int"), (64, 1000, 16, "int"), (300, 2048, 32, "scene"),
                                        (130, 4096, 64, "scene"), (50, 3000, 64, "int"), (9, 600, 200, "int"),
                                        (12, 5, 64, "normal"), (2048, 8192, 64, "scene")])
def test_knn_three_way(b200, oracle, refext, n, m, k, kind):
    q, r = clouds(n * 3 + k, 2, n, kind), clouds(m * 5 + k, 2, m, kind)
    d2_ref, idx_ref = refext.knn(k, q.cuda(), r.cuda())
    d2_o, idx_o = oracle.knn(k, q, r)
    assert torch.equal(idx_o, idx_ref.cpu()) and torch.equal(d2_o, d2_ref.cpu())
    d2, idx = b200.knn(k, q.cuda(), r.cuda())
    assert torch.equal(idx, idx_ref) and torch.equal(d2, d2_ref)
    # the operator layer's distance = torch.sqrt(dist2) on the GPU (reference pointnet2.py:103)
    dist, _ = b200.knn(k, q.cuda(), r.cuda(), sqrt=True)
    assert torch.equal(dist, torch.sqrt(d2_ref))


@pytest.mark.parametrize("n,m,kind", [(33, 1, "int"), (2048, 1024, "int"), (8192, 2048, "scene")])
def test_three_nn_three_way(b200, oracle, refext, n, m, kind):
    q, r = clouds(n, 2, n, kind), clouds(m + 1, 2, m, kind)
    d2_ref, idx_ref = refext.three_nn(q.cuda(), r.cuda())
    d2_o, idx_o = oracle.three_nn(q, r)
    assert torch.equal(idx_o, idx_ref.cpu()) and torch.equal(d2_o, d2_ref.cpu())
    d2, idx = b200.three_nn(q.cuda(), r.cuda())
    assert torch.equal(idx, idx_ref) and torch.equal(d2, d2_ref)


@pytest.mark.parametrize("n,m,r,ns,kind", [(100, 100, 2.0, 64, "int"), (1000, 300, 1.5, 16, "int"),
                                           (8192, 8192, 2.0, 64, "scene"), (9001, 100, 0.1, 8, "normal"),
                                           (64, 10, 1e-6, 5, "normal")])
def test_ball_query_three_way(b200, oracle, refext, n, m, r, ns, kind):
    xyz, c = clouds(n + 1, 2, n, kind), clouds(m + 2, 2, m, kind)
    ref = refext.ball_query(r, ns, xyz.cuda(), c.cuda())
    assert torch.equal(oracle.ball_query(r, ns, xyz, c), ref.cpu())
    assert torch.equal(b200.ball_query(r, ns, xyz.cuda(), c.cuda()), ref)


def test_group_gather_interpolate_three_way(b200, oracle, refext):
    rng = np.random.default_rng(9)
    B, C, N, M, S = 2, 12, 2048, 512, 32
    f = torch.from_numpy(rng.normal(size=(B, C, N)).astype(np.float32))
    idx = torch.from_numpy(rng.integers(0, N, size=(B, M, S)).astype(np.int32))
    ref = refext.group_points(f.cuda(), idx.cuda())
    assert torch.equal(oracle.group_points(f, idx), ref.cpu()) and torch.equal(b200.group_points(f.cuda(), idx.cuda()), ref)
    go = torch.from_numpy(rng.normal(size=(B, C, M, S)).astype(np.float32))
    gref = refext.group_points_grad(go.cuda(), idx.cuda(), N)
    torch.testing.assert_close(oracle.group_points_grad(go, idx, N), gref.cpu(), rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(b200.group_points_grad(go.cuda(), idx.cuda(), N), gref, rtol=1e-5, atol=1e-5)

    idx1 = idx[:, :, 0].contiguous()
    ref = refext.gather_points(f.cuda(), idx1.cuda())
    assert torch.equal(oracle.gather_points(f, idx1), ref.cpu()) and torch.equal(b200.gather_points(f.cuda(), idx1.cuda()), ref)

    n = 4096
    idx3 = torch.from_numpy(rng.integers(0, N, size=(B, n, 3)).astype(np.int32))
    w = torch.from_numpy(rng.random(size=(B, n, 3)).astype(np.float32))
    ref = refext.three_interpolate(f.cuda(), idx3.cuda(), w.cuda())
    assert torch.equal(oracle.three_interpolate(f, idx3, w), ref.cpu()), "interpolation rounding order"
    assert torch.equal(b200.three_interpolate(f.cuda(), idx3.cuda(), w.cuda()), ref)
    go3 = torch.from_numpy(rng.normal(size=(B, C, n)).astype(np.float32))
    gref = refext.three_interpolate_grad(go3.cuda(), idx3.cuda(), w.cuda(), N)
    torch.testing.assert_close(b200.three_interpolate_grad(go3.cuda(), idx3.cuda(), w.cuda(), N), gref, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(oracle.three_interpolate_grad(go3, idx3, w, N), gref.cpu(), rtol=1e-5, atol=1e-5)
