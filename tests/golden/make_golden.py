"""Generate the golden fixtures in tests/golden/ by running the UNMODIFIED reference Python
(/root/reference: models/segnet_*.py, losses/seg_loss_unsup.py, oa_icp.py) on CPU, on top of this
repo's operator layer bound to the CPU oracle.  Run in the build container only (the reference tree
does not travel to the GPU box); the produced .npz files are committed.

    python tests/golden/make_golden.py

Inputs are seeded; network weights come from `torch.manual_seed(10)` + this repo's MaskFormer3D
constructor and are loaded into the reference model through `load_state_dict` (identical parameter
names), so the fixture stores inputs + expected outputs but no weights.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.append("/root/reference")

from ogc_b200 import backend                      # noqa: E402
from oracle.pointnet2_oracle import OracleBackend  # noqa: E402
from tests.golden.cases import CASES, make_inputs, build_my_segnet   # noqa: E402


def main():
    backend.set_backend(OracleBackend())
    torch.set_num_threads(8)
    # the reference's MaskFormer head hard-codes .cuda() (utils/transformer_util.py:110)
    torch.Tensor.cuda = lambda self, *a, **k: self
    import importlib
    from losses import seg_loss_unsup as ref_loss

    for name, case in CASES.items():
        out = {}
        inp = make_inputs(case)
        if case["kind"] == "segnet":
            mine = build_my_segnet(case)
            ref_mod = importlib.import_module("models.segnet_%s" % case["variant"])
            ref = ref_mod.MaskFormer3D(n_slot=case["n_slot"], n_point=case["n_point"], use_xyz=True,
                                       n_transformer_layer=2, transformer_embed_dim=128,
                                       transformer_input_pos_enc=False)
            ref.load_state_dict(mine.state_dict())
            ref.train()
            pc = inp["pc"].clone()
            mask = ref(pc, pc)
            w = inp["probe"]
            (mask * w).sum().backward()
            out["mask"] = mask.detach().numpy()
            for pname in case["grad_params"]:
                out["grad:" + pname] = dict(ref.named_parameters())[pname].grad.numpy()
        elif case["kind"] == "ogc_loss":
            cfg = case["loss_cfg"]
            crit = ref_loss.UnsupervisedOGCLoss(
                ref_loss.DynamicLoss(**cfg["dynamic_loss_params"]), ref_loss.SmoothLoss(**cfg["smooth_loss_params"]),
                ref_loss.InvarianceLoss(**cfg["invariance_loss_params"]), ref_loss.EntropyLoss(), ref_loss.RankLoss(),
                weights=cfg["weights"], start_steps=cfg["start_steps"])
            logits = [l.clone().requires_grad_(True) for l in inp["logits"]]
            masks = [l.softmax(-1) for l in logits]
            loss, d = crit(inp["pcs"], masks, inp["flows"], step_w=True, it=case["it"], aug_transform=case["aug"])
            loss.backward()
            out["loss"] = np.float32(loss.item())
            for k, v in d.items():
                out["dict:" + k] = np.float32(v)
            for i, l in enumerate(logits):
                out["grad_logits%d" % i] = l.grad.numpy()
            R, t = ref_loss.fit_motion_svd_batch(
                inp["pcs"][0].unsqueeze(1).repeat(1, case["K"], 1, 1).reshape(-1, case["N"], 3),
                (inp["pcs"][0] + inp["flows"][0]).unsqueeze(1).repeat(1, case["K"], 1, 1).reshape(-1, case["N"], 3),
                masks[0].detach().transpose(1, 2).reshape(-1, case["N"]))
            out["kabsch_R"], out["kabsch_t"] = R.numpy(), t.numpy()
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, {k: getattr(v, "shape", v) for k, v in out.items()})


if __name__ == "__main__":
    main()
