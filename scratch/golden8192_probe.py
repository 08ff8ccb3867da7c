"""GPU: error statistics of the 8192-point goldens (decides the tolerances written into tests/test_golden.py)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests.golden.cases import CASES, make_inputs, build_my_segnet
from tests.test_golden import golden
from ogc_b200 import losses as L

name = "segnet_kitti_8192"; case = CASES[name]; inp = make_inputs(case); g = golden(name)
net = build_my_segnet(case).cuda(); pc = inp["pc"].cuda()
mask = net(pc, pc); (mask * inp["probe"].cuda()).sum().backward()
print("mask max abs err", float((mask.detach().cpu() - torch.from_numpy(g["mask"])).abs().max()))
for pname in case["grad_params"]:
    a = dict(net.named_parameters())[pname].grad.cpu().numpy(); b = g["grad:" + pname]
    d = np.abs(a - b); s = np.abs(b).max()
    print(pname, "fro-rel %.2e  max-rel %.2e  frac>1e-4*max %.4f  frac>1e-3*max %.5f" % (np.linalg.norm(a - b) / np.linalg.norm(b), d.max() / s, (d > 1e-4 * s).mean(), (d > 1e-3 * s).mean()))
name = "ogc_loss_8192_aug"; case = CASES[name]; inp = make_inputs(case); g = golden(name)
for fc in (True, False):
    crit = L.build_ogc_loss(case["loss_cfg"])
    logits = [l.clone().cuda().requires_grad_(True) for l in inp["logits"]]
    masks = [l.softmax(-1) for l in logits]
    L.FORCE_COMPOSED = fc
    loss, d = crit([p.cuda() for p in inp["pcs"]], masks, [f.cuda() for f in inp["flows"]], step_w=True, it=case["it"], aug_transform=True)
    loss.backward(); L.FORCE_COMPOSED = False
    print("composed" if fc else "fused", {k: (round(v, 6), round(float(g["dict:" + k]), 6)) for k, v in d.items()})
    for i, l in enumerate(logits):
        a = l.grad.cpu().numpy(); b = g["grad_logits%d" % i]; s = np.abs(b).max()
        print("  grad_logits%d fro-rel %.2e max-rel %.2e" % (i, np.linalg.norm(a - b) / np.linalg.norm(b), np.abs(a - b).max() / s))
