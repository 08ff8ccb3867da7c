"""Generate the golden fixtures in tests/golden/ by running the UNMODIFIED reference Python
(/root/reference: models/segnet_*.py, losses/seg_loss_unsup.py, oa_icp.py) on CPU, on top of this
repo's operator layer bound to the CPU oracle.  Run in the build container only (the reference tree
does not travel to the GPU box); the produced .npz files are committed.

    python tests/golden/make_golden.py [case names ...]

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

    only = set(sys.argv[1:])
    for name, case in CASES.items():
        if only and name not in only:
            continue
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
        elif case["kind"] == "flownet":
            from tests.golden.cases import build_my_flownet
            from losses import flow_loss_unsup as ref_floss
            mine = build_my_flownet(case)
            ref_mod = importlib.import_module("models.flownet_ogcdr")
            ref = ref_mod.FlowStep3D(npoint=case["npoint"], use_instance_norm=False, loc_flow_nn=8, loc_flow_rad=0.05)
            ref.load_state_dict(mine.state_dict())
            ref.train()
            preds = ref(inp["pc1"], inp["pc2"], inp["pc1"], inp["pc2"], iters=case["iters"])
            cfg = case["loss_cfg"]
            crit = ref_floss.UnsupervisedFlowStep3DLoss(ref_floss.ChamferLoss(**cfg["chamfer_loss_params"]),
                                                        ref_floss.SmoothLoss(**cfg["smooth_loss_params"]),
                                                        weights=cfg["weights"], iters_w=cfg["iters_w"])
            loss, d = crit(inp["pc1"], inp["pc2"], preds)
            loss.backward()
            for i, pr in enumerate(preds):
                out["flow%d" % i] = pr.detach().numpy()
            out["loss"] = np.float32(loss.item())
            for k, v in d.items():
                out["dict:" + k] = np.float32(v)
            for pname in case["grad_params"]:
                out["grad:" + pname] = dict(ref.named_parameters())[pname].grad.numpy().copy()
            # the same with 2 unrolled iterations: the GPU parity case (iteration 3 re-runs kNN on a cloud warped by
            # iteration 2's output and amplifies 1e-5 differences to 1e-2 -- seen between two fp32 CPU runs)
            ref.zero_grad()
            preds = ref(inp["pc1"], inp["pc2"], inp["pc1"], inp["pc2"], iters=2)
            crit.iters_w = cfg["iters_w"][:2]
            loss, d = crit(inp["pc1"], inp["pc2"], preds)
            loss.backward()
            out["i2:loss"] = np.float32(loss.item())
            for k, v in d.items():
                out["i2:dict:" + k] = np.float32(v)
            for pname in case["grad_params"]:
                out["i2:grad:" + pname] = dict(ref.named_parameters())[pname].grad.numpy().copy()
        elif case["kind"] == "oa_icp":
            # The reference's own oa_icp.object_aware_icp evaluated in FLOAT64 (its fp32 cdist is ill-conditioned:
            # SURVEY.md 7, hard part 8).  tensorboardX / metrics imports are stubbed; the k-NN index lookup inside
            # interpolate_mask_by_flow runs in fp32 (the operator layer is fp32-only), everything else in fp64.
            import types
            sys.modules.setdefault("tensorboardX", types.SimpleNamespace(SummaryWriter=object))
            import oa_icp as ref_icp
            orig_knn, orig_group = ref_loss.knn, ref_loss.grouping_operation
            ref_loss.knn = lambda k, a, b: tuple(t.double() if t.is_floating_point() else t for t in orig_knn(k, a.float().contiguous(), b.float().contiguous()))
            ref_loss.grouping_operation = lambda f, i: orig_group(f.float().contiguous(), i).double()
            orig_match = ref_icp.match_mask_by_iou
            ref_icp.match_mask_by_iou = lambda a, b: orig_match(a, b).double()
            torch.set_default_dtype(torch.float64)      # fit_motion_svd_batch builds its identity / zeros with the default dtype
            try:
                d = {k: v.double() for k, v in inp.items()}
                out["flow64"] = ref_icp.object_aware_icp(d["pc1"], d["pc2"], d["flow"], d["mask1"], d["mask2"],
                                                         icp_iter=case["icp_iter"]).numpy()
                out["kabsch_flow64"] = ref_icp.weighted_kabsch(d["pc1"], d["flow"], d["mask1"]).numpy()
            finally:
                torch.set_default_dtype(torch.float32)
                ref_icp.match_mask_by_iou = orig_match
                ref_loss.knn, ref_loss.grouping_operation = orig_knn, orig_group
            out["flow32"] = ref_icp.object_aware_icp(inp["pc1"], inp["pc2"], inp["flow"], inp["mask1"], inp["mask2"],
                                                     icp_iter=case["icp_iter"]).numpy()
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, {k: getattr(v, "shape", v) for k, v in out.items()})


if __name__ == "__main__":
    main()
