"""FlowStep3D -- host-side mirror of the reference scene-flow network (models/flownet_ogcdr.py:146-233 with the
building blocks of utils/flowstep3d_util.py:7-184) on the B200 operator set (BASELINE.json configs[2]).

Spec-driven: every set-abstraction block of the reference is one `FlowSA` instance created from a table, parameter
names are the reference's (`encoder_loc.sa1.mlp_convs.0.weight`, `...mlp_bns.0.running_mean`, `gru.convz...`,
`global_corr_layer.epsilon`, `flow_regressor.fc.weight`), so reference checkpoints load with load_state_dict.
On the GPU the blocks' shared MLPs (conv1x1 -> BatchNorm2d -> ReLU, x L, max over nsample) run fused through
ogc_b200/bn_fused.py (csrc/bn_mlp.cu + the pointwise contraction kernels); the torch expression below them is the CPU /
eval-mode path.  Differences (behaviour-preserving): furthest-point sampling of an unchanged cloud with an unchanged npoint is
memoised inside one forward pass (the reference recomputes the same FPS ~7 times per GRU iteration),
and the dead `knn=False` branch of FlowEmbedding (Appendix C.2) is not carried over.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

import pointnet2.pointnet2 as ops
from ogc_b200 import backend as _backend_mod
from ogc_b200 import bn_fused

USE_FUSED_MLP = True     # the blocks' (conv1x1, BatchNorm, ReLU) x L + max over nsample through csrc/bn_mlp.cu and the
                         # pointwise contraction kernels (ogc_b200/bn_fused.py); False: torch matmul + BatchNorm2d (tests)


def _fused_mlp_ok(x, convs, bns, use_act):
    """The fused block covers CUDA tensors on the b200 back-end, BatchNorm in training mode (batch statistics: the
    reference trains and evaluates its flow network that way only in train(); eval() takes the composed path)."""
    if not (USE_FUSED_MLP and x.is_cuda and x.dtype == torch.float32 and x.shape[0] > 0):
        return False
    if getattr(_backend_mod.get_backend(), "name", "") != "b200":
        return False
    if use_act and not all(bn.training and bn.affine for bn in bns):
        return False
    return bn_fused.supported([c.weight.shape[0] for c in convs], x.shape[-1])


def _conv1x1(conv, x):
    """The 1x1 convolution as an fp32 matmul on the conv's weight (cuDNN would pick TF32 kernels by default, which
    costs ~1e-3 of parity with the fp32 reference; the contraction is <= 214 channels wide)."""
    if not x.is_cuda:        # CPU arm (oracle-composed tests): the reference's own arithmetic, bit for bit
        return conv(x)
    w = conv.weight.view(conv.weight.shape[0], -1)
    return torch.matmul(w, x.flatten(2)).view(x.shape[0], w.shape[0], *x.shape[2:])


class _Bag(nn.Module):
    """Named container (gives sub-modules / parameters the reference's attribute paths)."""

    def __init__(self, **items):
        super().__init__()
        for k, v in items.items():
            setattr(self, k, v)


class FlowSA(nn.Module):
    """PointNetSetAbstraction (utils/flowstep3d_util.py:69-138): FPS -> kNN group (no radius) -> [xyz - centre ; feat]
    -> (conv1x1 -> BatchNorm2d -> ReLU) x L (or bare convs when use_act=False) -> max over nsample."""

    def __init__(self, npoint, nsample, in_channel, mlp, use_act=True):
        super().__init__()
        self.npoint, self.nsample, self.use_act = npoint, nsample, use_act
        self.mlp_convs, self.mlp_bns = nn.ModuleList(), nn.ModuleList()
        last = in_channel + 3
        for c in mlp:
            self.mlp_convs.append(nn.Conv2d(last, c, 1, bias=False))
            self.mlp_bns.append(nn.BatchNorm2d(c))
            last = c

    def forward(self, xyz, points, fps_idx=None, fps_cache=None):
        """xyz (B,3,N), points (B,D,N) -> new_xyz (B,3,M), new_points (B,C,M), fps_idx (B,M) int32."""
        xyz = xyz.contiguous()
        xyz_t = xyz.transpose(1, 2).contiguous()
        if fps_idx is None:
            # memo keyed by the tensor OBJECT (the entry holds a reference, so the id cannot be recycled while the
            # memo lives: one forward pass); the callers pass the same level tensors again and again
            key = (id(xyz), self.npoint)
            if fps_cache is not None and key in fps_cache and fps_cache[key][1] is xyz:
                fps_idx = fps_cache[key][0]
            else:
                fps_idx = ops.furthest_point_sample(xyz_t, self.npoint)
                if fps_cache is not None:
                    fps_cache[key] = (fps_idx, xyz)
        new_xyz = ops.gather_operation(xyz, fps_idx)
        new_xyz_t = new_xyz.transpose(1, 2).contiguous()
        _, idx = ops.knn(self.nsample, new_xyz_t, xyz_t)
        grouped = torch.cat([ops.grouping_operation(xyz, idx) - new_xyz.unsqueeze(-1),
                             ops.grouping_operation(points.contiguous(), idx)], dim=1)
        if _fused_mlp_ok(grouped, self.mlp_convs, self.mlp_bns, self.use_act):
            skip = 0 if xyz.requires_grad else 3        # the centred coordinates of a leaf cloud need no gradient
            return new_xyz, bn_fused.fused_bn_mlp(grouped, self.mlp_convs, self.mlp_bns if self.use_act else None, skip), fps_idx
        for conv, bn in zip(self.mlp_convs, self.mlp_bns):
            grouped = _conv1x1(conv, grouped)
            if self.use_act:
                grouped = F.relu(bn(grouped))
        return new_xyz, grouped.max(dim=-1).values, fps_idx


class FlowEmbedding(nn.Module):
    """utils/flowstep3d_util.py:7-66 (knn=True, corr_func='concat', max pooling)."""

    def __init__(self, radius, nsample, in_channel, mlp):
        super().__init__()
        self.radius, self.nsample = radius, nsample
        self.mlp_convs, self.mlp_bns = nn.ModuleList(), nn.ModuleList()
        last = in_channel * 2 + 3
        for c in mlp:
            self.mlp_convs.append(nn.Conv2d(last, c, 1, bias=False))
            self.mlp_bns.append(nn.BatchNorm2d(c))
            last = c

    def forward(self, pos1, pos2, feature1, feature2):
        pos1_t, pos2_t = pos1.transpose(1, 2).contiguous(), pos2.transpose(1, 2).contiguous()
        dist, idx = ops.knn(self.nsample, pos1_t, pos2_t)
        idx = ops.clip_neighbours_by_radius(dist, idx, self.radius)
        pos_diff = ops.grouping_operation(pos2.contiguous(), idx) - pos1.unsqueeze(-1)
        feat2 = ops.grouping_operation(feature2.contiguous(), idx)
        x = torch.cat([pos_diff, feat2, feature1.unsqueeze(-1).expand(-1, -1, -1, self.nsample)], dim=1)
        if _fused_mlp_ok(x, self.mlp_convs, self.mlp_bns, True):
            skip = 0 if (pos1.requires_grad or pos2.requires_grad) else 3
            return bn_fused.fused_bn_mlp(x, self.mlp_convs, self.mlp_bns, skip)
        for conv, bn in zip(self.mlp_convs, self.mlp_bns):
            x = F.relu(bn(_conv1x1(conv, x)))
        return x.max(dim=-1).values


def upsample_3nn(pos1, pos2, feature2):
    """PointNetFeaturePropogation with mlp=[] (utils/flowstep3d_util.py:155-175): inverse-distance 3-NN interpolation
    of feature2 (B,C,S) from pos2 (B,3,S) onto pos1 (B,3,N); distances floored at 1e-10."""
    dist, idx = ops.three_nn(pos1.transpose(1, 2).contiguous(), pos2.transpose(1, 2).contiguous())
    w = 1.0 / dist.clamp(min=1e-10)
    w = w / w.sum(dim=-1, keepdim=True)
    return (ops.grouping_operation(feature2.contiguous(), idx) * w.unsqueeze(1)).sum(dim=-1)


class FlowStep3D(nn.Module):
    def __init__(self, npoint=2048, use_instance_norm=False, loc_flow_nn=8, loc_flow_rad=0.1, k_decay_fact=1.0):
        super().__init__()
        if use_instance_norm:
            raise NotImplementedError("the reference configs use BatchNorm (config/flow/*/..yaml: use_instance_norm False)")
        n = npoint
        self.k_decay_fact = k_decay_fact
        sa = FlowSA
        self.encoder_loc = _Bag(sa1=sa(n // 2, 16, 3, [32, 32, 32]), sa2=sa(n // 4, 16, 32, [64, 64, 64]))
        self.encoder_glob = _Bag(sa1=sa(n // 8, 16, 64, [128, 128, 128]), sa2=sa(n // 16, 8, 128, [128, 128, 128]))
        self.global_corr_layer = _Bag(epsilon=nn.Parameter(torch.zeros(1)), sa1=sa(n // 8, 8, 3, [32, 64, 64]))
        self.h0_net = _Bag(sa1=sa(n // 4, 4, 64, [64, 64, 64]), sa2=sa(n // 4, 4, 64, [64], use_act=False))
        self.flow0_regressor = _Bag(sa1=sa(n // 4, 16, 64, [64, 64, 64]), fc=nn.Linear(64, 3))
        self.flow_regressor = _Bag(sa1=sa(n // 4, 16, 64, [64, 64, 64]), sa2=sa(n // 4, 16, 64, [64, 64, 64]),
                                   fc=nn.Linear(64, 3))
        self.local_corr_layer = FlowEmbedding(loc_flow_rad, loc_flow_nn, 64, [64, 64, 64])
        in_ch = 64 + (64 + 64 + 16 + 3)
        self.gru = _Bag(convz=sa(n // 4, 4, in_ch, [64], use_act=False), convr=sa(n // 4, 4, in_ch, [64], use_act=False),
                        convq=sa(n // 4, 4, in_ch, [64], use_act=False))
        self.flow_conv1 = sa(n // 4, 8, 3, [32, 32, 32])
        self.flow_conv2 = sa(n // 4, 4, 32, [16, 16, 16])

    # -- pieces, named after the reference methods -------------------------------------------------------------
    def _encode_loc(self, pc, feat, cache, fps_idx=None):
        a, b = self.encoder_loc.sa1, self.encoder_loc.sa2
        pc1, f1, i1 = a(pc, feat, None if fps_idx is None else fps_idx[0], cache)
        pc2, f2, i2 = b(pc1, f1, None if fps_idx is None else fps_idx[1], cache)
        return [pc, pc1, pc2], f2, [i1, i2]

    def _encode_glob(self, pc, feat, cache):
        pc1, f1, _ = self.encoder_glob.sa1(pc, feat, None, cache)
        pc2, f2, _ = self.encoder_glob.sa2(pc1, f1, None, cache)
        return [pc, pc1, pc2], f2

    def _global_corr(self, pc1_l, pc2_l, f1, f2, cache):
        """GlobalCorrLayer (models/flownet_ogcdr.py:40-76): soft correspondences on the coarsest level."""
        p1, p2 = pc1_l[2].transpose(1, 2), pc2_l[2].transpose(1, 2)
        g1, g2 = f1.transpose(1, 2), f2.transpose(1, 2)
        eps = torch.exp(self.global_corr_layer.epsilon) + 0.03
        d = (p1 ** 2).sum(-1, keepdim=True) + (p2 ** 2).sum(-1, keepdim=True).transpose(1, 2) - 2 * torch.bmm(p1, p2.transpose(1, 2))
        support = (d < 10 ** 2).float()
        g1 = g1 / torch.sqrt((g1 ** 2).sum(-1, keepdim=True) + 1e-8)
        g2 = g2 / torch.sqrt((g2 ** 2).sum(-1, keepdim=True) + 1e-8)
        corr = torch.exp(-(1.0 - torch.bmm(g1, g2.transpose(1, 2))) / eps) * support
        flow0 = (corr @ p2.contiguous()) / (corr.sum(-1, keepdim=True) + 1e-8) - p1.contiguous()
        flow0_us = upsample_3nn(pc1_l[1], pc1_l[2], flow0.transpose(1, 2).contiguous())
        _, feats_l1, _ = self.global_corr_layer.sa1(pc1_l[1], flow0_us, None, cache)
        return upsample_3nn(pc1_l[0], pc1_l[1], feats_l1)

    def _regress(self, bag, pc, feats, cache):
        x = feats
        for name in ("sa1", "sa2"):
            if hasattr(bag, name):
                _, x, _ = getattr(bag, name)(pc, x, None, cache)
        return bag.fc(x.transpose(1, 2)).transpose(1, 2).contiguous()

    def forward(self, pc1, pc2, feature1, feature2, iters=1):
        """pc*, feature* (B,N,3) -> list of `iters` flow predictions (B,N,3) (models/flownet_ogcdr.py:190-233)."""
        cache = {}
        pc1, pc2 = pc1.transpose(1, 2).contiguous(), pc2.transpose(1, 2).contiguous()
        f1, f2 = feature1.transpose(1, 2).contiguous(), feature2.transpose(1, 2).contiguous()
        pc1_l, feats1, fps1 = self._encode_loc(pc1, f1, cache)
        pc2_l, feats2, _ = self._encode_loc(pc2, f2, cache)
        g1_l, g1 = self._encode_glob(pc1_l[-1], feats1, cache)
        g2_l, g2 = self._encode_glob(pc2_l[-1], feats2, cache)
        corr_feats = self._global_corr(g1_l, g2_l, g1, g2, cache)
        flow0_lr = self._regress(self.flow0_regressor, pc1_l[2], corr_feats, cache)
        flow0 = upsample_3nn(pc1_l[0], pc1_l[2], flow0_lr)
        preds = [flow0.transpose(1, 2)]

        _, h, _ = self.h0_net.sa1(pc1_l[-1], feats1, None, cache)
        _, h, _ = self.h0_net.sa2(pc1_l[-1], h, None, cache)
        h = torch.tanh(h)
        pc1_new = pc1 + flow0.detach()
        pc1_new_lr = pc1_l[2] + flow0_lr.detach()
        lr_pc = pc1_l[2]
        for it in range(iters - 1):
            pc1_new, pc1_new_lr = pc1_new.detach(), pc1_new_lr.detach()
            flow_lr = pc1_new_lr - lr_pc
            new_l, feats1_new, _ = self._encode_loc(pc1_new, pc1_new, cache, fps1)
            corr = self.local_corr_layer(new_l[-1], pc2_l[-1], feats1_new, feats2)
            _, ff, _ = self.flow_conv1(lr_pc, flow_lr, None, cache)
            _, ff, _ = self.flow_conv2(lr_pc, ff, None, cache)
            x = torch.cat([feats1_new, corr, ff, flow_lr], dim=1)
            hx = torch.cat([h, x], dim=1)
            z = torch.sigmoid(self.gru.convz(lr_pc, hx, None, cache)[1])
            r = torch.sigmoid(self.gru.convr(lr_pc, hx, None, cache)[1])
            q = torch.tanh(self.gru.convq(lr_pc, torch.cat([r * h, x], dim=1), None, cache)[1])
            h = (1 - z) * h + z * q
            delta_lr = self._regress(self.flow_regressor, lr_pc, h, cache) / (self.k_decay_fact * it + 1)
            pc1_new_lr = pc1_new_lr + delta_lr
            pc1_new = pc1_new + upsample_3nn(pc1_l[0], lr_pc, delta_lr)
            preds.append((pc1_new - pc1).transpose(1, 2))
        return preds


# ------------------------------------------------------------------------------------------------------------------
# Unsupervised scene-flow loss                                                  losses/flow_loss_unsup.py:7-140
# ------------------------------------------------------------------------------------------------------------------
class ChamferLoss(nn.Module):
    def __init__(self, loss_norm=2):
        super().__init__()
        self.loss_norm = loss_norm

    def forward(self, pc1, pc2, flow):
        warped = (pc1 + flow).contiguous()
        pc2 = pc2.contiguous()
        w_t, p2_t = warped.transpose(1, 2).contiguous(), pc2.transpose(1, 2).contiguous()
        _, idx = ops.knn(1, warped, pc2)
        d1 = (w_t - ops.grouping_operation(p2_t, idx.detach()).squeeze(-1)).norm(p=self.loss_norm, dim=1)
        _, idx = ops.knn(1, pc2, warped)
        d2 = (p2_t - ops.grouping_operation(w_t, idx.detach()).squeeze(-1)).norm(p=self.loss_norm, dim=1)
        return (d1 + d2).mean()


class FlowSmoothLoss(nn.Module):
    """SmoothLoss of losses/flow_loss_unsup.py:92-109: w_knn * kNN term + w_ball_q * ball-query term on the flow."""

    def __init__(self, w_knn, w_ball_q, knn_loss_params, ball_q_loss_params):
        super().__init__()
        self.w_knn, self.w_ball_q = w_knn, w_ball_q
        self.knn, self.ball = dict(knn_loss_params), dict(ball_q_loss_params)

    @staticmethod
    def _term(flow_t, idx, p):
        return (flow_t.unsqueeze(3) - ops.grouping_operation(flow_t, idx.detach())).norm(p=p, dim=1).mean()

    def forward(self, pc, flow):
        pc = pc.contiguous()
        flow_t = flow.transpose(1, 2).contiguous()
        dist, idx = ops.knn(self.knn["k"], pc, pc)
        idx = ops.clip_neighbours_by_radius(dist, idx, self.knn["radius"])
        l_knn = self._term(flow_t, idx, self.knn.get("loss_norm", 1))
        bidx = ops.ball_query(self.ball["radius"], self.ball["k"], pc, pc)
        l_ball = self._term(flow_t, bidx, self.ball.get("loss_norm", 1))
        return self.w_knn * l_knn + self.w_ball_q * l_ball


class UnsupervisedFlowStep3DLoss(nn.Module):
    def __init__(self, chamfer_loss, smooth_loss, weights=[0.75, 0.25], iters_w=[1.0]):
        super().__init__()
        self.chamfer_loss, self.smooth_loss = chamfer_loss, smooth_loss
        self.w_chamfer, self.w_smooth = weights
        self.iters_w = iters_w
        self.defer_logging = False       # True: return the logged scalars as ONE device tensor (no host sync: capturable)

    def forward(self, pc1, pc2, flow_preds):
        assert len(flow_preds) == len(self.iters_w)
        terms, logged = [], {}
        for i, flow in enumerate(flow_preds):
            c = self.chamfer_loss(pc1, pc2, flow)
            s = self.smooth_loss(pc1, flow)
            logged[f"chamfer_loss_#{i}"], logged[f"smooth_loss_#{i}"] = c, s
            terms.append(self.iters_w[i] * (self.w_chamfer * c + self.w_smooth * s))
        loss = sum(terms)
        logged["sum"] = loss
        keys = list(logged)
        stacked = torch.stack([logged[k].detach().float().reshape(()) for k in keys])
        if self.defer_logging:
            return loss, {"_keys": keys, "_values": stacked}
        return loss, dict(zip(keys, stacked.tolist()))                                        # one D2H, not 2*iters+1


OGCDR_FLOW_LOSS_CFG = {   # config/flow/ogcdr/ogcdr_unsup.yaml:37-52
    "weights": [0.75, 0.25], "iters_w": [0.5, 0.3, 0.3, 0.3],
    "chamfer_loss_params": {"loss_norm": 2},
    "smooth_loss_params": {"w_knn": 3.0, "w_ball_q": 1.0,
                           "knn_loss_params": {"k": 4, "radius": 0.05, "loss_norm": 1},
                           "ball_q_loss_params": {"k": 8, "radius": 0.1, "loss_norm": 1}},
}


def build_flow_loss(cfg=OGCDR_FLOW_LOSS_CFG):
    return UnsupervisedFlowStep3DLoss(ChamferLoss(**cfg["chamfer_loss_params"]), FlowSmoothLoss(**cfg["smooth_loss_params"]),
                                      weights=cfg["weights"], iters_w=cfg["iters_w"])
