"""Golden-vector tests.  tests/golden/*.npz were produced by the UNMODIFIED reference Python
(models/segnet_*.py, losses/seg_loss_unsup.py) on CPU over this repo's operator layer + CPU oracle
(tests/golden/make_golden.py).  Here this repo's own network / loss mirrors must reproduce them:
  * on CPU through the oracle back-end (composed implementation)        -- not gpu
  * on a B200 through libogc_b200 (composed AND fused implementations)  -- gpu
Tolerance: 1e-4 absolute on masks / losses (BASELINE.json north_star), 1e-4 relative-to-max on
gradients.
"""
import os

import numpy as np
import pytest
import torch

from tests.golden.cases import CASES, make_inputs, build_my_segnet

HERE = os.path.dirname(os.path.abspath(__file__))


def golden(name):
    return dict(np.load(os.path.join(HERE, "golden", name + ".npz")))


def assert_close_rel_to_max(got, ref, tol, what):
    scale = max(float(np.abs(ref).max()), 1e-12)
    err = float(np.abs(got - ref).max())
    assert err <= tol * scale, f"{what}: max abs err {err:.3e} vs scale {scale:.3e}"


def run_segnet_case(name, device, grad_tol=2e-4, fro_tol=None):
    case = CASES[name]
    inp = make_inputs(case)
    net = build_my_segnet(case).to(device)
    pc = inp["pc"].to(device)
    mask = net(pc, pc)
    (mask * inp["probe"].to(device)).sum().backward()
    g = golden(name)
    err = float((mask.detach().cpu() - torch.from_numpy(g["mask"])).abs().max())
    assert err <= 1e-4, f"{name}: mask max abs err {err:.3e}"
    params = dict(net.named_parameters())
    for pname in case["grad_params"]:
        got, ref = params[pname].grad.cpu().numpy(), g["grad:" + pname]
        if fro_tol is not None:
            rel = float(np.linalg.norm(got - ref) / max(np.linalg.norm(ref), 1e-30))
            assert rel <= fro_tol, f"{name} grad {pname}: Frobenius-relative error {rel:.2e}"
        assert_close_rel_to_max(got, ref, grad_tol, f"{name} grad {pname}")


def run_loss_case(name, device, force_composed, grad_tol=2e-4):
    from ogc_b200 import losses as L
    case = CASES[name]
    inp = make_inputs(case)
    g = golden(name)
    crit = L.build_ogc_loss(case["loss_cfg"])
    logits = [l.clone().to(device).requires_grad_(True) for l in inp["logits"]]
    masks = [l.softmax(-1) for l in logits]
    pcs = [p.to(device) for p in inp["pcs"]]
    flows = [f.to(device) for f in inp["flows"]]
    L.FORCE_COMPOSED = force_composed
    try:
        loss, d = crit(pcs, masks, flows, step_w=True, it=case["it"], aug_transform=case["aug"])
        loss.backward()
    finally:
        L.FORCE_COMPOSED = False
    assert abs(loss.item() - float(g["loss"])) <= 1e-4 * max(1.0, abs(float(g["loss"])))
    for k in ["dynamic", "smooth", "invariance", "entropy", "sum"]:
        assert abs(d[k] - float(g["dict:" + k])) <= 1e-4 * max(1.0, abs(float(g["dict:" + k]))), k
    assert abs(d["rank"] - float(g["dict:rank"])) <= 1e-4 * abs(float(g["dict:rank"])), "rank"
    for i, l in enumerate(logits):
        assert_close_rel_to_max(l.grad.cpu().numpy(), g["grad_logits%d" % i], grad_tol, f"{name} grad_logits{i}")
    return pcs, flows, masks, g


# ------------------------------------------------------------------------------- CPU (oracle)
@pytest.mark.parametrize("name", ["segnet_sapien_512", "segnet_kitti_1024"])
def test_segnet_matches_reference_cpu(oracle_ops, name):
    torch.set_num_threads(8)
    run_segnet_case(name, "cpu")


@pytest.mark.parametrize("name", ["ogc_loss_aug", "ogc_loss_noaug"])
def test_ogc_loss_matches_reference_cpu(oracle_ops, name):
    from ogc_b200 import losses as L
    pcs, flows, masks, g = run_loss_case(name, "cpu", force_composed=True)
    case = CASES[name]
    rep = lambda x: x.unsqueeze(1).repeat(1, case["K"], 1, 1).reshape(-1, case["N"], 3)
    R, t = L.fit_motion_svd_batch(rep(pcs[0]), rep(pcs[0] + flows[0]), masks[0].detach().transpose(1, 2).reshape(-1, case["N"]))
    np.testing.assert_allclose(R.numpy(), g["kabsch_R"], atol=2e-5)
    np.testing.assert_allclose(t.numpy(), g["kabsch_t"], atol=2e-4)


def run_segnet_8192(device):
    """The headline size: 2 clouds x 8192 points, n_slot 10, golden = the unmodified reference on CPU.
    Masks: 1e-4 absolute (measured on a B200: 1.7e-6).  Gradients of everything downstream of the set-abstraction stack
    (FP, transformer, object MLP: no data-dependent decisions) to 1e-4 of the tensor maximum (measured 1e-6 .. 2.5e-6).
    Set-abstraction weight gradients additionally absorb DECISION FLIPS: a ReLU input or a max-pool margin within an
    ulp of zero resolves differently under a different fp32 summation order, and the flipped path's O(1) contribution
    spreads over every upstream weight -- measured 4e-4 between two CPU fp32 runs that differ only in conv2d vs matmul,
    1e-3 .. 2e-3 Frobenius on the GPU.  tests/test_gpu_fused_sa.py::test_gradients_match_fp64_given_identical_decisions
    pins the kernels' ARITHMETIC to <= 1e-4 with the decisions held fixed and counts the flips, so the 5e-3 here is
    about flips only."""
    case = CASES["segnet_kitti_8192"]
    inp = make_inputs(case)
    net = build_my_segnet(case).to(device)
    pc = inp["pc"].to(device)
    mask = net(pc, pc)
    (mask * inp["probe"].to(device)).sum().backward()
    g = golden("segnet_kitti_8192")
    err = float((mask.detach().cpu() - torch.from_numpy(g["mask"])).abs().max())
    assert err <= 1e-4, f"mask max abs err {err:.3e}"
    params = dict(net.named_parameters())
    for pname in case["grad_params"]:
        got, ref = params[pname].grad.cpu().numpy(), g["grad:" + pname]
        rel = float(np.linalg.norm(got - ref) / np.linalg.norm(ref))
        if pname.startswith("SA_modules"):
            assert rel <= 5e-3, f"{pname}: Frobenius-relative error {rel:.2e}"
        else:
            assert rel <= 1e-4, f"{pname}: Frobenius-relative error {rel:.2e}"
            assert_close_rel_to_max(got, ref, 1e-4, pname)


def test_segnet_8192_matches_reference_cpu(oracle_ops):
    """BASELINE.json configs[1] at its own size through the composed path + CPU oracle."""
    torch.set_num_threads(8)
    run_segnet_8192("cpu")


def test_ogc_loss_8192_matches_reference_cpu(oracle_ops):
    torch.set_num_threads(8)
    run_loss_case("ogc_loss_8192_aug", "cpu", force_composed=True)


# ------------------------------------------------------------------------------- GPU (B200)
@pytest.mark.gpu
def test_segnet_8192_matches_reference_gpu(b200):
    run_segnet_8192("cuda")


@pytest.mark.gpu
@pytest.mark.parametrize("force_composed", [True, False], ids=["composed", "fused"])
def test_ogc_loss_8192_matches_reference_gpu(b200, force_composed):
    """4 views x 8192 points, K = 10, all loss terms: loss dict to 1e-4, d loss / d logits to 1e-3 of the maximum
    (measured 1.8e-4: the smoothness terms sum |m_n - m_nbr| over 96 neighbours per point, whose sign decisions sit at
    exact ties for duplicated neighbours) and 1e-4 in the Frobenius norm (measured 2.5e-5)."""
    from ogc_b200 import losses as L
    case = CASES["ogc_loss_8192_aug"]
    inp = make_inputs(case)
    g = golden("ogc_loss_8192_aug")
    crit = L.build_ogc_loss(case["loss_cfg"])
    logits = [l.clone().cuda().requires_grad_(True) for l in inp["logits"]]
    masks = [l.softmax(-1) for l in logits]
    L.FORCE_COMPOSED = force_composed
    try:
        loss, d = crit([p.cuda() for p in inp["pcs"]], masks, [f.cuda() for f in inp["flows"]], step_w=True,
                       it=case["it"], aug_transform=True)
        loss.backward()
    finally:
        L.FORCE_COMPOSED = False
    for k in ["dynamic", "smooth", "invariance", "entropy", "rank", "sum"]:
        ref = float(g["dict:" + k])
        assert abs(d[k] - ref) <= 1e-4 * max(1.0, abs(ref)), (k, d[k], ref)
    for i, l in enumerate(logits):
        got, ref = l.grad.cpu().numpy(), g["grad_logits%d" % i]
        assert float(np.linalg.norm(got - ref) / np.linalg.norm(ref)) <= 1e-4, i
        assert_close_rel_to_max(got, ref, 1e-3, f"grad_logits{i}")



@pytest.mark.gpu
@pytest.mark.parametrize("name", ["segnet_sapien_512", "segnet_kitti_1024"])
def test_segnet_matches_reference_gpu(b200, name):
    # masks to 1e-4 abs.  Weight gradients pass through softmax(cos/0.05) and atomically-ordered fp32 sums:
    # the golden (CPU fp32) and the GPU differ by summation order alone at the 5e-4 level relative to the
    # largest entry (the reference cannot reproduce its own gradients bit-for-bit either);
    # a single flipped ReLU / arg-max decision moves one channel's gradient by a few 1e-3 of the tensor max
    # (analysed in tests/test_gpu_fused_sa.py).
    run_segnet_case(name, "cuda", grad_tol=5e-3)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["ogc_loss_aug", "ogc_loss_noaug"])
@pytest.mark.parametrize("force_composed", [True, False], ids=["composed", "fused"])
def test_ogc_loss_matches_reference_gpu(b200, name, force_composed):
    pcs, flows, masks, g = run_loss_case(name, "cuda", force_composed)
    case = CASES[name]
    Rt = b200.weighted_kabsch(pcs[0], flows[0], masks[0].detach().contiguous(), second_is_flow=True)
    R = Rt[..., :9].reshape(-1, 3, 3).cpu().numpy()
    t = Rt[..., 9:].reshape(-1, 3).cpu().numpy()
    np.testing.assert_allclose(R, g["kabsch_R"], atol=2e-5)
    np.testing.assert_allclose(t, g["kabsch_t"], atol=2e-4)
