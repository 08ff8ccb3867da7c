"""GPU: the uniform-grid neighbourhood kernels (csrc/grid.cu) must return EXACTLY what the brute-force kernels return
(which are bit-exact against the reference extension: tests/test_gpu_vs_refext.py): bounded k-NN under the
(distance, index) order of interpolate_gpu.cu:41-51, ball query under the first-nsample-in-index-order rule of
ball_query_gpu.cu:27-44 -- on KITTI-SF-like scenes, uniform clouds, integer lattices full of exact ties, degenerate
(flat / single-point / tiny-radius) clouds and separate query clouds."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _clouds(kind, B, N, seed):
    rng = np.random.default_rng(seed)
    if kind == "kitti":
        from ogc_b200 import data
        return torch.cat([data.make_batch(seed + i, 1, N, aug=False)[0][:, 0] for i in range(B)]).cuda().contiguous()
    if kind == "lattice":            # exact distance ties everywhere
        return torch.from_numpy(rng.integers(-6, 7, size=(B, N, 3)).astype(np.float32)).cuda()
    if kind == "flat":
        p = rng.uniform(-20, 20, size=(B, N, 3)).astype(np.float32)
        p[..., 1] = 0.25
        return torch.from_numpy(p).cuda()
    if kind == "point":
        return torch.full((B, N, 3), 1.5, device="cuda")
    return torch.from_numpy((rng.uniform(-1, 1, size=(B, N, 3)) * np.array([30, 3, 20])).astype(np.float32)).cuda()


def _both(fn):
    from ogc_b200.backend import B200Backend
    out = {}
    for grid in (False, True):
        B200Backend.USE_GRID = grid
        try:
            out[grid] = fn()
        finally:
            B200Backend.USE_GRID = True
    torch.cuda.synchronize()
    return out[False], out[True]


@pytest.mark.parametrize("kind,N,k,radius", [("kitti", 8192, 32, 1.0), ("kitti", 8192, 64, 2.0), ("uniform", 4096, 16, 1.5),
                                              ("lattice", 2048, 32, 2.0), ("lattice", 3000, 8, 1.0), ("flat", 5000, 32, 0.7),
                                              ("point", 1500, 8, 0.5), ("uniform", 2048, 40, 0.01), ("uniform", 1100, 100, 50.0)])
def test_grid_knn_equals_brute_force(b200, kind, N, k, radius):
    pc = _clouds(kind, 3, N, 11)
    (d0, i0), (d1, i1) = _both(lambda: b200.knn_bounded(k, pc, pc, radius))
    assert torch.equal(i0, i1), float((i0 != i1).float().mean())
    assert torch.equal(d0, d1)


@pytest.mark.parametrize("kind,N,ns,radius", [("kitti", 8192, 64, 2.0), ("uniform", 4096, 16, 1.0), ("lattice", 2048, 32, 2.0),
                                               ("lattice", 3000, 8, 1.0), ("flat", 5000, 64, 1.2), ("point", 1500, 8, 0.5),
                                               ("uniform", 2048, 16, 0.01), ("uniform", 1100, 64, 50.0)])
def test_grid_ball_query_equals_brute_force(b200, kind, N, ns, radius):
    pc = _clouds(kind, 3, N, 12)
    a, b = _both(lambda: b200.ball_query(radius, ns, pc, pc))
    assert torch.equal(a, b), float((a != b).float().mean())


def test_grid_with_a_separate_query_cloud(b200):
    """QueryAndGroup's use: centres (a subset moved slightly, some far outside the box) against the full cloud."""
    pc = _clouds("kitti", 2, 8192, 5)
    q = pc[:, ::4].clone()
    q[:, :7] += 100.0                      # queries outside the bounding box: no neighbour within the radius
    q[:, 7:40] += 0.013
    q = q.contiguous()
    (d0, i0), (d1, i1) = _both(lambda: b200.knn_bounded(64, q, pc, 1.2))
    assert torch.equal(i0, i1) and torch.equal(d0, d1)
    a, b = _both(lambda: b200.ball_query(1.2, 32, pc, q))
    assert torch.equal(a, b)
