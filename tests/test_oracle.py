"""CPU tests of the oracle (oracle/pointnet2_oracle.c) against brute-force definitions.

The reference ships no golden vectors, so the restatement is checked against
(a) the behavioural spec of SURVEY.md Appendix A written as straightforward numpy, on
integer-valued coordinates where every fp32 operation is exact and ties are everywhere, and
(b) an independent, literal Python simulation of the reference's FPS block reduction.
"""
import numpy as np
import pytest
import torch


def int_cloud(rng, b, n, lo=-4, hi=5):
    """Small-integer coordinates: all distance arithmetic is exact in fp32 -> ties are exact."""
    return torch.from_numpy(rng.integers(lo, hi, size=(b, n, 3)).astype(np.float32))


def d2_exact(a, b):
    a = a.astype(np.float64)[:, None, :]
    b = b.astype(np.float64)[None, :, :]
    return ((a - b) ** 2).sum(-1)


# ---------------------------------------------------------------------------------------------
def fps_definition(pts, m, bs):
    """Appendix A: argmax of running min distance; ties -> smallest (bitrev(k mod bs), k)."""
    n = len(pts)
    nbits = int(np.log2(bs))
    def bitrev(t):
        return int(format(t, f"0{nbits}b")[::-1], 2) if nbits else 0
    key = np.array([(bitrev(k % bs), k) for k in range(n)])
    order = np.lexsort((key[:, 1], key[:, 0]))          # candidates in tie-priority order
    mind = np.full(n, 1e10)
    sel = [0]
    for _ in range(1, m):
        d = ((pts.astype(np.float64) - pts[sel[-1]].astype(np.float64)) ** 2).sum(-1)
        mind = np.minimum(mind, d)
        best = mind.max()
        cands = order[mind[order] == best]
        sel.append(int(cands[0]))
    return np.array(sel, dtype=np.int32)


def fps_literal_tree(pts, m, bs):
    """Independent literal simulation of sampling_gpu.cu:93-209 (per-thread scan + shared-memory tree)."""
    n = len(pts)
    p = pts.astype(np.float64)
    temp = np.full(n, 1e10)
    sel = [0]
    for _ in range(1, m):
        old = sel[-1]
        dists = np.full(bs, -1.0)
        dists_i = np.zeros(bs, dtype=np.int64)
        for tid in range(bs):
            best, besti = -1.0, 0
            for k in range(tid, n, bs):
                d = ((p[k] - p[old]) ** 2).sum()
                d2 = min(d, temp[k])
                temp[k] = d2
                if d2 > best:
                    best, besti = d2, k
            dists[tid], dists_i[tid] = best, besti
        half = bs // 2
        while half >= 1:
            for tid in range(half):
                v1, v2 = dists[tid], dists[tid + half]
                i1, i2 = dists_i[tid], dists_i[tid + half]
                dists[tid] = max(v1, v2)
                dists_i[tid] = i2 if v2 > v1 else i1
            half //= 2
        sel.append(int(dists_i[0]))
    return np.array(sel, dtype=np.int32)


@pytest.mark.parametrize("n,m", [(1, 1), (2, 2), (3, 3), (5, 4), (16, 16), (37, 20), (64, 64), (100, 37), (130, 130)])
def test_fps_ties_integer_grid(oracle, n, m):
    from oracle import pointnet2_oracle as O
    rng = np.random.default_rng(n * 131 + m)
    xyz = int_cloud(rng, 3, n, -2, 3)  # heavy duplication -> many exact ties
    got = oracle.fps(xyz, m).numpy()
    bs = O.opt_n_threads(n)
    assert bs == min(1024, 2 ** int(np.floor(np.log2(n))))
    for b in range(3):
        np.testing.assert_array_equal(got[b], fps_definition(xyz[b].numpy(), m, bs))
        np.testing.assert_array_equal(got[b], fps_literal_tree(xyz[b].numpy(), m, bs))


def test_fps_random_float_matches_definition(oracle):
    rng = np.random.default_rng(7)
    xyz = torch.from_numpy(rng.normal(size=(2, 1500, 3)).astype(np.float32))
    got = oracle.fps(xyz, 64).numpy()
    for b in range(2):
        # continuous data: no ties, and float64 distances order like fp32 ones except at 1-ulp gaps
        ref = fps_definition(xyz[b].numpy(), 64, 1024)
        assert (got[b] == ref).mean() > 0.95
    assert got[:, 0].tolist() == [0, 0]
    assert all(len(set(r.tolist())) == 64 for r in got)


def test_fps_of_fps_prefix_is_identity(oracle):
    """FPS on an FPS-ordered prefix returns 0..M-1 (FlowStep3D relies on it)."""
    rng = np.random.default_rng(3)
    xyz = torch.from_numpy(rng.normal(size=(1, 700, 3)).astype(np.float32))
    order = oracle.fps(xyz, 256).long()
    sub = xyz[0][order[0]].unsqueeze(0).contiguous()
    again = oracle.fps(sub, 256)
    np.testing.assert_array_equal(again[0].numpy(), np.arange(256, dtype=np.int32))
    # and M == N is legal
    full = oracle.fps(sub, sub.shape[1])
    np.testing.assert_array_equal(full[0].numpy(), np.arange(256, dtype=np.int32))


# ---------------------------------------------------------------------------------------------
def knn_definition(q, ref, k):
    d = d2_exact(q, ref)
    m = ref.shape[0]
    idx = np.zeros((len(q), k), dtype=np.int32)
    dist = np.full((len(q), k), np.inf, dtype=np.float32)
    for i in range(len(q)):
        order = np.lexsort((np.arange(m), d[i]))[:k]      # (distance, index) lexicographic
        idx[i, :len(order)] = order
        dist[i, :len(order)] = d[i, order]
    return dist, idx


@pytest.mark.parametrize("n,m,k", [(17, 50, 1), (33, 64, 3), (9, 200, 16), (40, 300, 32), (12, 500, 64),
                                   (5, 10, 16), (3, 1, 4), (4, 260, 200)])
def test_knn_ties_integer_grid(oracle, n, m, k):
    rng = np.random.default_rng(n + 7 * m + 13 * k)
    q, ref = int_cloud(rng, 2, n), int_cloud(rng, 2, m)
    d2, idx = oracle.knn(k, q, ref)
    for b in range(2):
        rd, ri = knn_definition(q[b].numpy(), ref[b].numpy(), k)
        np.testing.assert_array_equal(idx[b].numpy(), ri)
        np.testing.assert_array_equal(d2[b].numpy(), rd)   # includes the (+inf, 0) tail when m < k


def test_three_nn_equals_knn3(oracle):
    rng = np.random.default_rng(11)
    q, ref = int_cloud(rng, 2, 70), int_cloud(rng, 2, 90)
    a = oracle.three_nn(q, ref)
    b = oracle.knn(3, q, ref)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    q2, ref2 = int_cloud(rng, 1, 5), int_cloud(rng, 1, 2)  # m < 3
    a = oracle.three_nn(q2, ref2)
    b = oracle.knn(3, q2, ref2)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    assert torch.isinf(a[0][..., 2]).all() and (a[1][..., 2] == 0).all()


def test_knn_ignores_nan_and_inf_candidates(oracle):
    q = torch.zeros(1, 1, 3)
    ref = torch.tensor([[[1., 0, 0], [float("nan"), 0, 0], [float("inf"), 0, 0], [2., 0, 0]]])
    d2, idx = oracle.knn(4, q, ref)
    assert idx[0, 0].tolist() == [0, 3, 0, 0]
    assert d2[0, 0, :2].tolist() == [1.0, 4.0] and torch.isinf(d2[0, 0, 2:]).all()


# ---------------------------------------------------------------------------------------------
def ball_query_definition(xyz, centres, r, ns):
    d = d2_exact(centres, xyz)
    r2 = float(np.float32(r) * np.float32(r))
    out = np.zeros((len(centres), ns), dtype=np.int32)
    for i in range(len(centres)):
        hits = np.nonzero(d[i] < r2)[0][:ns]
        if len(hits):
            out[i, :] = hits[0]
            out[i, :len(hits)] = hits
    return out


@pytest.mark.parametrize("n,m,r,ns", [(100, 30, 2.0, 8), (300, 40, 3.0, 64), (50, 10, 0.5, 4), (64, 64, 100.0, 16)])
def test_ball_query_integer_grid(oracle, n, m, r, ns):
    rng = np.random.default_rng(n + m)
    xyz, c = int_cloud(rng, 2, n), int_cloud(rng, 2, m)
    got = oracle.ball_query(r, ns, xyz, c)
    for b in range(2):
        np.testing.assert_array_equal(got[b].numpy(), ball_query_definition(xyz[b].numpy(), c[b].numpy(), r, ns))


def test_ball_query_radius_is_strict(oracle):
    xyz = torch.tensor([[[2., 0, 0], [1., 0, 0]]])
    c = torch.zeros(1, 1, 3)
    assert oracle.ball_query(2.0, 2, xyz, c)[0, 0].tolist() == [1, 1]     # d2 == r2 is NOT a hit
    assert oracle.ball_query(0.5, 2, xyz, c)[0, 0].tolist() == [0, 0]     # no hit -> zeros


# ---------------------------------------------------------------------------------------------
def test_group_gather_interpolate_forward_backward(oracle):
    rng = np.random.default_rng(5)
    B, C, N, M, S = 2, 5, 40, 7, 6
    f = torch.from_numpy(rng.normal(size=(B, C, N)).astype(np.float32))
    idx = torch.from_numpy(rng.integers(0, N, size=(B, M, S)).astype(np.int32))
    out = oracle.group_points(f, idx)
    ref = torch.gather(f.unsqueeze(2).expand(B, C, M, N), 3, idx.long().unsqueeze(1).expand(B, C, M, S))
    assert torch.equal(out, ref)
    go = torch.from_numpy(rng.normal(size=(B, C, M, S)).astype(np.float32))
    g = oracle.group_points_grad(go, idx, N)
    gref = torch.zeros(B, C, N).scatter_add_(2, idx.long().view(B, 1, M * S).expand(B, C, M * S), go.view(B, C, M * S))
    torch.testing.assert_close(g, gref, rtol=1e-5, atol=1e-6)

    idx1 = idx[:, :, 0].contiguous()
    assert torch.equal(oracle.gather_points(f, idx1), torch.gather(f, 2, idx1.long().unsqueeze(1).expand(B, C, M)))
    go1 = go[..., 0].contiguous()
    g1 = oracle.gather_points_grad(go1, idx1, N)
    g1ref = torch.zeros(B, C, N).scatter_add_(2, idx1.long().unsqueeze(1).expand(B, C, M), go1)
    torch.testing.assert_close(g1, g1ref, rtol=1e-5, atol=1e-6)

    n = 9
    idx3 = torch.from_numpy(rng.integers(0, N, size=(B, n, 3)).astype(np.int32))
    w = torch.from_numpy(rng.random(size=(B, n, 3)).astype(np.float32))
    o3 = oracle.three_interpolate(f, idx3, w)
    picked = torch.gather(f.unsqueeze(2).expand(B, C, n, N), 3, idx3.long().unsqueeze(1).expand(B, C, n, 3))
    torch.testing.assert_close(o3, (picked.double() * w.double().unsqueeze(1)).sum(-1).float(), rtol=1e-6, atol=1e-6)
    go3 = torch.from_numpy(rng.normal(size=(B, C, n)).astype(np.float32))
    g3 = oracle.three_interpolate_grad(go3, idx3, w, N)
    contrib = (go3.unsqueeze(-1) * w.unsqueeze(1)).reshape(B, C, n * 3)
    g3ref = torch.zeros(B, C, N).scatter_add_(2, idx3.long().view(B, 1, n * 3).expand(B, C, n * 3), contrib)
    torch.testing.assert_close(g3, g3ref, rtol=1e-5, atol=1e-6)


def test_three_interpolate_rounding_order(oracle):
    """out = fma(w2,p2, fma(w0,p0, w1*p1)) -- pick values where another order differs in the last bit."""
    f = torch.tensor([[[1.0000001, 3.0000002, -4.0000005]]])
    idx = torch.tensor([[[0, 1, 2]]], dtype=torch.int32)
    w = torch.tensor([[[0.3333333, 0.3333334, 0.3333333]]])
    got = oracle.three_interpolate(f, idx, w).item()
    p = [float(np.float32(v)) for v in f[0, 0].tolist()]
    ww = [float(np.float32(v)) for v in w[0, 0].tolist()]
    from fractions import Fraction as Fr
    t = float(np.float32(ww[1] * p[1]))                       # fp32 product (double product is exact, one rounding)
    t = float(np.float32(float(Fr(ww[0]) * Fr(p[0]) + Fr(t))))  # fma: exact, then one rounding (via double: see note)
    t = float(np.float32(float(Fr(ww[2]) * Fr(p[2]) + Fr(t))))
    # float(Fraction) rounds to double first; the extra rounding can differ only in astronomically rare
    # half-way cases, not for these constants.
    assert got == t


@pytest.mark.parametrize("B,cin,M,S,widths,use_act", [(2, 6, 12, 4, [16, 16], True), (3, 35, 10, 5, [32, 16, 16], True),
                                                       (2, 22, 9, 4, [16], False)])
def test_flow_mlp_oracle_matches_torch_autograd(B, cin, M, S, widths, use_act):
    """oracle/flow_mlp_oracle.py (the BatchNorm shared MLP of the FlowStep3D blocks, utils/flowstep3d_util.py:126-137, in
    the statistics / coefficient form csrc/bn_mlp.cu evaluates) against torch's own modules and autograd in fp64."""
    import torch.nn as nn
    import torch.nn.functional as F
    from oracle import flow_mlp_oracle as O
    g = torch.Generator().manual_seed(B * 100 + cin)
    x = torch.randn(B, cin, M, S, generator=g, dtype=torch.float64, requires_grad=True)
    probe = torch.randn(B, widths[-1], M, generator=g, dtype=torch.float64)
    convs, bns, last = [], [], cin
    for c in widths:
        convs.append(nn.Conv2d(last, c, 1, bias=False).double())
        bn = nn.BatchNorm2d(c).double()
        with torch.no_grad():
            bn.weight.add_(0.3 * torch.randn(c, generator=g, dtype=torch.float64))
            bn.bias.add_(0.2 * torch.randn(c, generator=g, dtype=torch.float64))
        bns.append(bn)
        last = c
    running = [(bn.running_mean.clone().numpy(), bn.running_var.clone().numpy()) for bn in bns]
    h = x
    for conv, bn in zip(convs, bns):
        h = conv(h)
        if use_act:
            h = F.relu(bn(h))
    ref = h.max(dim=-1).values
    (ref * probe).sum().backward()

    Ws = [c.weight.detach().numpy().reshape(c.weight.shape[0], -1) for c in convs]
    gam = [b.weight.detach().numpy() for b in bns] if use_act else None
    bet = [b.bias.detach().numpy() for b in bns] if use_act else None
    out, cache = O.forward(x.detach().numpy(), Ws, gam, bet, running if use_act else None)
    dx, dWs, dgs, dbs = O.backward(probe.numpy(), cache)
    close = lambda a, b: np.allclose(a, b, rtol=1e-9, atol=1e-11)
    assert close(out, ref.detach().numpy())
    assert close(dx, x.grad.numpy())
    for l, c in enumerate(convs):
        assert close(dWs[l], c.weight.grad.numpy().reshape(dWs[l].shape))
    if use_act:
        for l, b in enumerate(bns):
            assert close(dgs[l], b.weight.grad.numpy()) and close(dbs[l], b.bias.grad.numpy())
            assert close(running[l][0], b.running_mean.numpy()) and close(running[l][1], b.running_var.numpy())
