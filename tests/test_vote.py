"""Multi-frame mask voting (reference vote.py:17-131).  Oracle = the reference's own
`mask_voting` evaluated in float64 on CPU (tests/golden/vote.npz, make_golden_vote.py); parity bound 1e-4 with the fp32
reference's own deviation recorded beside it.  The mirror never builds an N x N correspondence: chained
correspondences are applied as repeated softmax-weighted transfers (ogc_b200/vote.py)."""
import os

import numpy as np
import pytest
import torch

from tests.golden.cases import CASES, make_inputs

HERE = os.path.dirname(os.path.abspath(__file__))
G = dict(np.load(os.path.join(HERE, "golden", "vote.npz")))


def run(device):
    from ogc_b200 import vote
    case = CASES["vote"]
    inp = {k: v.to(device) for k, v in make_inputs(case).items()}
    return vote.mask_voting(inp["pc"], inp["mask"], inp["flows"], time_window_size=case["window"]).cpu().numpy()


def check(voted):
    ref_dev = float(np.abs(G["voted32"] - G["voted64"]).max())
    err = float(np.abs(voted - G["voted64"]).max())
    assert err <= 1e-4, f"mask_voting: {err:.2e} from the fp64 reference (fp32 reference itself: {ref_dev:.2e})"
    np.testing.assert_allclose(voted.sum(-1), 1.0, atol=1e-5)


def test_mask_voting_composed_cpu():
    check(run("cpu"))


def test_chained_transfer_equals_propagated_correspondence():
    """The identity the mirror rests on, in float64 with the reference's formulas: normalise(C01 @ C12) @ M ==
    C01 @ (C12 @ M)."""
    from ogc_b200 import vote
    torch.manual_seed(0)
    pc = torch.randn(3, 200, 3, dtype=torch.float64)
    fl = torch.randn(2, 200, 3, dtype=torch.float64) * 0.05
    M = torch.softmax(torch.randn(1, 200, 6, dtype=torch.float64), -1)
    c01 = vote.pairwise_correspondence(pc[0:1], pc[1:2], fl[0:1])
    c12 = vote.pairwise_correspondence(pc[1:2], pc[2:3], fl[1:2])
    c02 = torch.bmm(c01, c12)
    c02 = c02 / c02.sum(-1, keepdim=True).clamp(1e-10)              # vote.py:54-55
    chained = vote.transfer(pc[0:1] + fl[0:1], pc[1:2], vote.transfer(pc[1:2] + fl[1:2], pc[2:3], M))
    assert float((torch.bmm(c02, M) - chained).abs().max()) < 1e-12


@pytest.mark.gpu
def test_mask_voting_fused_gpu(b200):
    check(run("cuda"))


@pytest.mark.gpu
def test_softmax_transfer_matches_dense_formulation(b200):
    """The streaming kernel against the dense softmax(-cdist/T) @ V in float64, ragged sizes, K = 1..16."""
    torch.manual_seed(1)
    for B, N1, N2, K in ((2, 700, 900, 10), (1, 33, 1500, 1), (3, 513, 512, 16)):
        q = torch.randn(B, N1, 3, device="cuda") * 2
        key = torch.cat([q[:, :min(N1, N2)] + 0.01 * torch.randn(B, min(N1, N2), 3, device="cuda"),
                         torch.randn(B, max(N2 - N1, 0), 3, device="cuda") * 2], 1)[:, :N2].contiguous()
        val = torch.softmax(torch.randn(B, N2, K, device="cuda") * 2, -1)
        got = b200.softmax_transfer(q, key, val, 0.01)
        corr = (-torch.cdist(q.double(), key.double(), compute_mode="donot_use_mm_for_euclid_dist") / 0.01).softmax(-1)
        ref = torch.bmm(corr, val.double())
        assert float((got.double() - ref).abs().max()) < 1e-4, (B, N1, N2, K)


@pytest.mark.gpu
def test_mask_voting_full_size_properties(b200):
    """KITTI-SF size (N = 8192, K = 10, 4 frames): the reference would need 268 MB per correspondence and N^3 bmms;
    here: rows stay normalised, a static scene with zero flow and identical frames votes every mask onto itself."""
    from ogc_b200 import vote
    torch.manual_seed(2)
    T, N, K = 4, 8192, 10
    # a jittered 32 x 8 x 32 lattice (spacing 1 m): every other point is >= 0.8 m away, exp(-80) of the weight of a point itself
    g = torch.stack(torch.meshgrid(torch.arange(32.0), torch.arange(8.0), torch.arange(32.0), indexing="ij"), -1).reshape(1, N, 3)
    pc0 = (g + (torch.rand(1, N, 3) - 0.5) * 0.2).cuda()[:, torch.randperm(N)]
    pc = pc0.expand(T, -1, -1).contiguous()
    mask0 = torch.softmax(torch.randn(1, N, K, device="cuda") * 3, -1)
    mask = mask0.expand(T, -1, -1).contiguous()
    flows = torch.zeros(T - 1, 2, N, 3, device="cuda")
    voted = vote.mask_voting(pc, mask, flows, time_window_size=3)
    assert voted.shape == (T, N, K)
    torch.testing.assert_close(voted.sum(-1), torch.ones(T, N, device="cuda"), atol=1e-5, rtol=0)
    torch.testing.assert_close(voted, mask, atol=1e-4, rtol=0)
