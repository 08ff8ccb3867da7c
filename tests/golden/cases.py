"""Seeded definitions of the golden cases (shared by make_golden.py and the tests)."""
import numpy as np
import torch

KITTI_LOSS = {
    "weights": [10.0, 0.1, 0.1], "start_steps": [0, 100, 1000],
    "dynamic_loss_params": {"loss_norm": 2},
    "smooth_loss_params": {"w_knn": 3.0, "w_ball_q": 1.0,
                           "knn_loss_params": {"k": 32, "radius": 1.0, "loss_norm": 1},
                           "ball_q_loss_params": {"k": 64, "radius": 2.0, "loss_norm": 1}},
    "invariance_loss_params": {"loss_norm": 2},
}
SAPIEN_LOSS = {   # config/seg/sapien/sapien_unsup.yaml
    "weights": [10.0, 0.1, 0.1], "start_steps": [0, 0, 0],
    "dynamic_loss_params": {"loss_norm": 2},
    "smooth_loss_params": {"w_knn": 3.0, "w_ball_q": 1.0,
                           "knn_loss_params": {"k": 8, "radius": 0.1, "loss_norm": 1},
                           "ball_q_loss_params": {"k": 16, "radius": 0.2, "loss_norm": 1}},
    "invariance_loss_params": {"loss_norm": 2},
}

CASES = {
    # BASELINE.json configs[0]: SAPIEN 512-pt cloud through segnet_sapien (+ a probe backward)
    "segnet_sapien_512": {"kind": "segnet", "variant": "sapien", "n_slot": 8, "n_point": 512, "B": 2, "seed": 10,
                          "scale": 0.5,
                          "grad_params": ["SA_modules.0.mlps.0.layer0.conv.weight", "FP_modules.0.mlp.layer2.conv.weight",
                                          "MF_head.query.weight"]},
    # a reduced KITTI-shaped net (n_point 1024 -> 256/128/64 centres), metre-scale scene
    "segnet_kitti_1024": {"kind": "segnet", "variant": "kitti", "n_slot": 10, "n_point": 1024, "B": 2, "seed": 11,
                          "scale": 12.0,
                          "grad_params": ["SA_modules.0.mlps.1.layer2.conv.weight", "SA_modules.2.mlps.0.layer0.conv.weight",
                                          "object_mlp.1.conv.bias"]},
    "ogc_loss_aug": {"kind": "ogc_loss", "B": 2, "N": 640, "K": 6, "seed": 12, "aug": True, "it": 5000,
                     "scale": 6.0, "loss_cfg": KITTI_LOSS},
    "ogc_loss_noaug": {"kind": "ogc_loss", "B": 3, "N": 512, "K": 8, "seed": 13, "aug": False, "it": 50,
                       "scale": 0.4, "loss_cfg": SAPIEN_LOSS},
}


# BASELINE.json configs[1] at its OWN size: one KITTI-SF-like pair (2 clouds x 8192 points, the
# synthetic road scene of ogc_b200/data.py), n_slot 10, through the unmodified reference on CPU
CASES["segnet_kitti_8192"] = {"kind": "segnet", "variant": "kitti", "n_slot": 10, "n_point": 8192, "B": 2, "seed": 31,
                              "scene": "kittisf",
                              "grad_params": ["SA_modules.0.mlps.0.layer0.conv.weight", "SA_modules.1.mlps.0.layer1.conv.weight",
                                              "SA_modules.2.mlps.0.layer2.conv.weight", "FP_modules.0.mlp.layer1.conv.weight",
                                              "MF_head.transformer_layers.1.cross_attn.in_proj_weight", "object_mlp.1.conv.bias"]}
CASES["ogc_loss_8192_aug"] = {"kind": "ogc_loss", "B": 1, "N": 8192, "K": 10, "seed": 32, "aug": True, "it": 100000,
                              "scene": "kittisf", "loss_cfg": KITTI_LOSS}


FLOW_LOSS = {   # config/flow/ogcdr/ogcdr_unsup.yaml:37-52 with 3 unrolled iterations
    "weights": [0.75, 0.25], "iters_w": [0.5, 0.3, 0.3],
    "chamfer_loss_params": {"loss_norm": 2},
    "smooth_loss_params": {"w_knn": 3.0, "w_ball_q": 1.0,
                           "knn_loss_params": {"k": 4, "radius": 0.05, "loss_norm": 1},
                           "ball_q_loss_params": {"k": 8, "radius": 0.1, "loss_norm": 1}},
}
CASES["flownet_ogcdr_512"] = {"kind": "flownet", "npoint": 512, "B": 2, "seed": 15, "iters": 3, "loss_cfg": FLOW_LOSS,
                              "grad_params": ["encoder_loc.sa1.mlp_convs.0.weight", "gru.convq.mlp_convs.0.weight",
                                              "flow_regressor.fc.weight", "global_corr_layer.epsilon"]}
# BASELINE configs[2]'s own cloud size (ogcdr: 2048 points), one pair per sample, 2 + 3 unrolled iterations as above
CASES["flownet_ogcdr_2048"] = {"kind": "flownet", "npoint": 2048, "B": 2, "seed": 16, "iters": 3, "loss_cfg": FLOW_LOSS,
                               "grad_params": ["encoder_loc.sa1.mlp_convs.0.weight", "encoder_glob.sa1.mlp_bns.1.weight",
                                               "local_corr_layer.mlp_convs.1.weight", "gru.convq.mlp_convs.0.weight",
                                               "flow_regressor.sa1.mlp_convs.2.weight", "flow_regressor.fc.weight",
                                               "global_corr_layer.epsilon"]}
CASES["oa_icp"] = {"kind": "oa_icp", "B": 2, "N": 768, "K": 6, "seed": 14, "scale": 8.0, "icp_iter": 4}
CASES["vote"] = {"kind": "vote", "T": 5, "N": 384, "K": 5, "seed": 21, "window": 3}


def make_inputs(case):
    rng = np.random.default_rng(case["seed"])
    f32 = lambda a: torch.from_numpy(np.asarray(a, dtype=np.float32))
    if case["kind"] == "vote":
        # T frames of one scene: K rigid parts moving by small per-frame motions, points re-sampled (permuted + jittered)
        # per frame, soft masks with a different slot order per frame, noisy adjacent flows in both directions
        T, N, K = case["T"], case["N"], case["K"]
        base = rng.uniform(-1, 1, size=(N, 3))
        seg = np.clip(((base[:, 0] + 1) / 2 * K).astype(int), 0, K - 1)
        frames, segs = [], []
        cur = base.copy()
        for t in range(T):
            perm = rng.permutation(N)
            frames.append(cur[perm] + rng.normal(size=(N, 3)) * 0.002)
            segs.append(seg[perm])
            nxt = cur.copy()
            for k in range(K):
                ang = rng.uniform(-0.05, 0.05)
                Rz = np.array([[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1]])
                nxt[seg == k] = cur[seg == k] @ Rz.T + rng.uniform(-0.05, 0.05, size=3)
            cur = nxt
        pc = np.stack(frames)
        flows = np.zeros((T - 1, 2, N, 3))
        for t in range(T - 1):
            for a, b, slot in ((t, t + 1, 0), (t + 1, t, 1)):
                d = ((pc[a][:, None, :] - pc[b][None, :, :]) ** 2).sum(-1)
                # flow towards a nearby point of the same part in the other frame, plus noise
                d = d + 1e3 * (segs[a][:, None] != segs[b][None, :])
                flows[t, slot] = pc[b][d.argmin(1)] - pc[a] + rng.normal(size=(N, 3)) * 0.01
        masks = []
        for t in range(T):
            lg = rng.normal(size=(N, K)) * 0.8
            lg[np.arange(N), segs[t]] += 2.5
            e = np.exp(lg - lg.max(-1, keepdims=True))
            masks.append((e / e.sum(-1, keepdims=True))[:, rng.permutation(K)])
        return {"pc": f32(pc), "mask": f32(np.stack(masks)), "flows": f32(flows)}
    if case["kind"] == "segnet" and case.get("scene") == "kittisf":
        from ogc_b200 import data
        pcs = data.make_batch(case["seed"], case["B"] // 2, case["n_point"], aug=False)[0]          # (b,2,N,3)
        pc = pcs.reshape(case["B"], case["n_point"], 3).clone()
        probe = f32(rng.normal(size=(case["B"], case["n_point"], case["n_slot"])))
        return {"pc": pc, "probe": probe}
    if case["kind"] == "ogc_loss" and case.get("scene") == "kittisf":
        from ogc_b200 import data
        pcs, segms, flows, _ = data.make_batch(case["seed"], case["B"], case["N"], aug=case["aug"])   # (b,V,N,3)
        V, K = pcs.shape[1], case["K"]
        logits = []
        for v in range(V):
            lg = rng.normal(size=(case["B"], case["N"], K)) * 0.5
            seg = segms[:, v].numpy() % K
            lg[np.arange(case["B"])[:, None], np.arange(case["N"])[None, :], seg] += 2.0
            logits.append(f32(lg))
        return {"pcs": [pcs[:, v].clone() for v in range(V)], "flows": [flows[:, v].clone() for v in range(V)], "logits": logits}
    if case["kind"] == "segnet":
        pc = f32(rng.uniform(-1, 1, size=(case["B"], case["n_point"], 3)) * case["scale"])
        probe = f32(rng.normal(size=(case["B"], case["n_point"], case["n_slot"])))
        return {"pc": pc, "probe": probe}
    if case["kind"] == "flownet":
        B, N = case["B"], case["npoint"]
        pc1 = rng.uniform(-0.5, 0.5, size=(B, N, 3))
        ang = 0.05
        Rz = np.array([[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1]])
        pc2 = np.stack([(pc1[b] @ Rz.T + np.array([0.03, -0.02, 0.01]))[rng.permutation(N)] for b in range(B)])
        pc2 = pc2 + rng.normal(size=pc2.shape) * 0.002
        return {"pc1": f32(pc1), "pc2": f32(pc2)}
    if case["kind"] == "oa_icp":
        B, N, K = case["B"], case["N"], case["K"]
        pc1 = rng.uniform(-1, 1, size=(B, N, 3)) * case["scale"]
        seg = np.clip(((pc1[..., 0] / case["scale"] + 1) / 2 * K).astype(int), 0, K - 1)
        true_flow = np.zeros_like(pc1)
        for k in range(K):
            ang = rng.uniform(-0.08, 0.08)
            Rz = np.array([[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1]])
            t = rng.uniform(-0.3, 0.3, size=3)
            sel = seg == k
            true_flow[sel] = pc1[sel] @ Rz.T + t - pc1[sel]
        pc2 = np.stack([(pc1[b] + true_flow[b])[rng.permutation(N)] for b in range(B)])
        pc2 = pc2 + rng.normal(size=pc2.shape) * 0.01
        flow0 = true_flow + rng.normal(size=pc1.shape) * 0.15
        def soft(segm, noise):
            lg = rng.normal(size=(B, N, K)) * noise
            lg[np.arange(B)[:, None], np.arange(N)[None, :], segm] += 3.0
            e = np.exp(lg - lg.max(-1, keepdims=True))
            return e / e.sum(-1, keepdims=True)
        d = ((pc2[:, :, None, :] - (pc1 + true_flow)[:, None, :, :]) ** 2).sum(-1)
        seg2 = np.take_along_axis(seg, d.argmin(-1), 1)
        perm_slots = rng.permutation(K)
        mask2 = soft(seg2, 0.5)[..., perm_slots]          # frame 2 uses another slot order
        return {"pc1": f32(pc1), "pc2": f32(pc2), "flow": f32(flow0), "mask1": f32(soft(seg, 0.5)), "mask2": f32(mask2)}
    V = 4 if case["aug"] else 2
    B, N, K = case["B"], case["N"], case["K"]
    pcs, flows, logits = [], [], []
    for v in range(V):
        pc = rng.uniform(-1, 1, size=(B, N, 3)) * case["scale"]
        # piecewise-rigid flow: K blobs along x, each with its own small rotation + translation
        seg = np.clip(((pc[..., 0] / case["scale"] + 1) / 2 * K).astype(int), 0, K - 1)
        flow = np.zeros_like(pc)
        for k in range(K):
            ang = rng.uniform(-0.1, 0.1)
            Rz = np.array([[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1]])
            t = rng.uniform(-0.05, 0.05, size=3) * case["scale"]
            sel = seg == k
            flow[sel] = pc[sel] @ Rz.T + t - pc[sel]
        flow += rng.normal(size=flow.shape) * 0.002 * case["scale"]
        lg = rng.normal(size=(B, N, K)) * 0.5
        lg[np.arange(B)[:, None], np.arange(N)[None, :], seg] += 2.0
        pcs.append(f32(pc)); flows.append(f32(flow)); logits.append(f32(lg))
    return {"pcs": pcs, "flows": flows, "logits": logits}


def build_my_segnet(case):
    from ogc_b200.segnet import MaskFormer3D
    torch.manual_seed(10)
    net = MaskFormer3D(n_slot=case["n_slot"], n_point=case["n_point"], variant=case["variant"])
    # non-trivial norm parameters (fresh GroupNorm / LayerNorm are identity-affine)
    g = torch.Generator().manual_seed(case["seed"])
    with torch.no_grad():
        for n, p in net.named_parameters():
            if n.endswith("gn.weight") or (("norm" in n) and n.endswith("weight")):
                p.add_(0.2 * torch.randn(p.shape, generator=g))
            elif n.endswith("gn.bias") or (("norm" in n) and n.endswith("bias")):
                p.add_(0.1 * torch.randn(p.shape, generator=g))
    net.train()
    return net


def build_my_flownet(case):
    from ogc_b200.flownet import FlowStep3D
    torch.manual_seed(10)
    net = FlowStep3D(npoint=case["npoint"], loc_flow_nn=8, loc_flow_rad=0.05)
    g = torch.Generator().manual_seed(case["seed"])
    with torch.no_grad():
        for n, p in net.named_parameters():
            if "mlp_bns" in n and n.endswith("weight"):
                p.add_(0.2 * torch.randn(p.shape, generator=g))
            elif "mlp_bns" in n and n.endswith("bias"):
                p.add_(0.1 * torch.randn(p.shape, generator=g))
    net.train()
    return net
