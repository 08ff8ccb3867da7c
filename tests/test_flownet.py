"""FlowStep3D + unsupervised flow loss (BASELINE.json configs[2]; reference models/flownet_ogcdr.py:146-233,
losses/flow_loss_unsup.py:7-140).  Golden = the unmodified reference network and loss on CPU with OUR state_dict
loaded (tests/golden/make_golden.py), so the parameter naming is part of what is checked.

Two goldens: 3 unrolled iterations (CPU arm: same arithmetic as the reference, checked to 1e-5) and 2 iterations (GPU
arm, 1e-4).  The third iteration re-runs kNN / radius clipping on the cloud warped by the second one's output and
amplifies 1e-5 differences to ~1e-2 in an untrained BatchNorm network (measured between two fp32 CPU runs that only
differ in the conv algorithm), so on the GPU it is only required to stay close in the median."""
import os

import numpy as np
import pytest
import torch

from tests.golden.cases import CASES, build_my_flownet, make_inputs

HERE = os.path.dirname(os.path.abspath(__file__))
G = dict(np.load(os.path.join(HERE, "golden", "flownet_ogcdr_512.npz")))
CASE = CASES["flownet_ogcdr_512"]
G2048 = dict(np.load(os.path.join(HERE, "golden", "flownet_ogcdr_2048.npz")))     # BASELINE configs[2]'s cloud size
CASE2048 = CASES["flownet_ogcdr_2048"]


def run(device, iters, CASE=CASE):
    from ogc_b200.flownet import build_flow_loss
    net = build_my_flownet(CASE).to(device)
    inp = {k: v.to(device) for k, v in make_inputs(CASE).items()}
    preds = net(inp["pc1"], inp["pc2"], inp["pc1"], inp["pc2"], iters=iters)
    cfg = dict(CASE["loss_cfg"], iters_w=CASE["loss_cfg"]["iters_w"][:iters])
    loss, d = build_flow_loss(cfg)(inp["pc1"], inp["pc2"], preds)
    loss.backward()
    grads = {n: dict(net.named_parameters())[n].grad.cpu().numpy() for n in CASE["grad_params"]}
    return [p.detach().cpu().numpy() for p in preds], float(loss.detach()), d, grads


def check(preds, loss, d, grads, prefix, tol, gtol, G=G):
    for i, p in enumerate(preds):
        err = float(np.abs(p - G["flow%d" % i]).max())
        print(f"flow prediction {i}: max |diff| {err:.2e} (|flow| up to {float(np.abs(G['flow%d' % i]).max()):.2f})")
        assert err <= tol, f"flow prediction {i}: {err:.2e}"
    ref = float(G[prefix + "loss"])
    assert abs(loss - ref) <= 10 * tol * max(1.0, abs(ref)), (loss, ref)
    for k, v in d.items():
        ref = float(G[prefix + "dict:" + k])
        assert abs(v - ref) <= 10 * tol * max(1.0, abs(ref)), (k, v, ref)
    for n, g in grads.items():
        ref = G[prefix + "grad:" + n]
        rel = float(np.linalg.norm(g - ref) / max(np.linalg.norm(ref), 1e-12))
        print(f"grad {n}: rel Frobenius error {rel:.2e}")
        assert rel <= gtol, f"grad {n}: rel Frobenius error {rel:.2e}"


def test_flownet_state_dict_names_match_reference_layout():
    names = set(build_my_flownet(CASE).state_dict())
    for n in CASE["grad_params"] + ["encoder_glob.sa2.mlp_bns.2.running_var", "h0_net.sa2.mlp_convs.0.weight",
                                    "local_corr_layer.mlp_convs.0.weight", "flow_conv2.mlp_bns.0.num_batches_tracked",
                                    "gru.convz.mlp_convs.0.weight", "flow0_regressor.fc.bias"]:
        assert n in names, n


def test_flownet_composed_cpu_matches_reference(oracle_ops):
    check(*run("cpu", 3), prefix="", tol=1e-5, gtol=1e-4)
    check(*run("cpu", 2), prefix="i2:", tol=1e-5, gtol=1e-4)
    check(*run("cpu", 2, CASE2048), prefix="i2:", tol=1e-5, gtol=1e-4, G=G2048)      # configs[2]'s own cloud size


@pytest.mark.gpu
def test_flownet_gpu_matches_reference(b200):
    check(*run("cuda", 2), prefix="i2:", tol=1e-4, gtol=5e-3)     # measured 5e-4
    preds = run("cuda", 3)[0]
    med = float(np.median(np.abs(preds[2] - G["flow2"])))
    print(f"third iteration: median |diff| {med:.2e}")
    assert np.isfinite(preds[2]).all() and med < 2e-2


@pytest.mark.gpu
def test_flownet_gpu_matches_reference_at_2048_points(b200):
    """The same at the cloud size BASELINE configs[2] is quoted on (golden from the unmodified reference on CPU)."""
    check(*run("cuda", 2, CASE2048), prefix="i2:", tol=1e-4, gtol=5e-3, G=G2048)


@pytest.mark.gpu
def test_flow_trainer_graph_replay_equals_eager(b200):
    """ogc_b200.train.FlowTrainer: the CUDA-graph replay of the step (train_flow.py:59-92) must train like the eager
    step -- same logged losses, same parameters after 3 steps (BatchNorm running statistics included), up to the fp32
    summation order of the atomically accumulated gradients (which differs from run to run in the reference too)."""
    from ogc_b200.flownet import FlowStep3D, build_flow_loss, OGCDR_FLOW_LOSS_CFG
    from ogc_b200.train import FlowTrainer

    def run(graphed):
        torch.manual_seed(10)
        net = FlowStep3D(npoint=512, loc_flow_nn=8, loc_flow_rad=0.05).cuda()
        tr = FlowTrainer(net, build_flow_loss(dict(OGCDR_FLOW_LOSS_CFG, iters_w=[0.5, 0.3, 0.3])), iters=3)
        g = torch.Generator().manual_seed(4)
        pc1 = torch.rand(4, 512, 3, generator=g) - 0.5
        pcs = torch.stack([pc1, pc1 + 0.02 + 0.003 * torch.randn(4, 512, 3, generator=g)], 1)
        batch = (pcs.pin_memory(), None, None, None)
        logs = [(tr.train_step_graphed if graphed else tr.train_step)(batch) for _ in range(3)]
        torch.cuda.synchronize()
        bn = torch.cat([m.running_var.flatten() for m in net.modules() if isinstance(m, torch.nn.BatchNorm2d)])
        return logs, tr.opt.flat_p.clone(), bn.clone()

    le, pe, be_ = run(False)
    le2, pe2, be2 = run(False)
    le3 = run(False)[0]
    lg, pg, bg = run(True)
    # step 0 starts from identical parameters: the logged losses must agree to fp32 summation order
    for k in le[0]:
        assert abs(le[0][k] - lg[0][k]) <= 2e-5 * max(1.0, abs(le[0][k])), (k, le[0][k], lg[0][k])
    # later steps: Adam's first updates are sign-like (lr * g / (|g| + eps)), so gradient entries at the noise level of
    # the atomically ordered sums move their parameter by +-lr in either run; the yardstick is therefore the deviation
    # between two EAGER runs, not zero
    dev = lambda la, lb: max(abs(a[k] - b[k]) / max(1.0, abs(a[k])) for a, b in zip(la, lb) for k in a)
    noise_l, noise_p = max(dev(le, le2), dev(le, le3), dev(le2, le3)), float((pe - pe2).abs().max())
    noise_b = float((be_ - be2).abs().max()) / float(be_.abs().max())
    print(f"eager vs eager: logs {noise_l:.2e} params {noise_p:.2e} bn {noise_b:.2e}; "
          f"graph vs eager: logs {dev(le, lg):.2e} params {float((pe - pg).abs().max()):.2e} "
          f"bn {float((be_ - bg).abs().max()) / float(be_.abs().max()):.2e}")
    # steps 1-2 run on parameters that already differ by such +-lr moves, and the third unrolled iteration amplifies them
    # (module docstring: 1e-5 -> 1e-2; eager-vs-eager spreads of 5e-3 ... 4e-2 were measured on a B200): a floor of 5e-2 on
    # top of the measured eager-vs-eager spread (a replay that does not train, or trains on stale buffers, is off by 0.5)
    assert dev(le, lg) <= 5 * noise_l + 5e-2
    assert float((pe - pg).abs().max()) <= 5 * noise_p + 1e-4
    assert float((pe - pg).abs().mean()) <= 5 * float((pe - pe2).abs().mean()) + 1e-6
    assert float((be_ - bg).abs().max()) / float(be_.abs().max()) <= 5 * noise_b + 1e-4
