"""MaskFormer3D -- host-side mirror of the reference segmentation networks
(models/segnet_kitti.py:12-89, models/segnet_sapien.py, models/segnet_ogcdr.py) built on the
B200 operator set.

Not a copy: one spec-driven implementation covers the three reference variants; 1x1 convolutions
are evaluated as fp32 matmuls (cuBLAS sgemm -- no TF32, independent of torch.backends.cudnn flags);
the duplicate k-NN of the multi-scale SA level is computed once.
Parameter names and shapes are identical to the reference (e.g.
`SA_modules.0.mlps.1.layer2.conv.weight (64,32,1,1)`, `...normlayer.gn.weight`,
`MF_head.transformer_layers.0.cross_attn.in_proj_weight`), so reference checkpoints load with
`load_state_dict` and the parity tests move weights both ways.

Reference call stack reproduced:
  SA level   utils/pointnet2_util.py:16-49   FPS -> gather centres -> per scale [kNN(64) -> radius clip
             -> group xyz,feat -> centre -> concat -> (conv1x1, GroupNorm(4), ReLU) x3 -> max over nsample]
  FP level   utils/pointnet2_util.py:96-120  three_nn -> inverse-distance weights -> interpolate ->
             concat skip -> (conv1x1, GN, ReLU) xL
  head       utils/transformer_util.py:62-121, models/segnet_kitti.py:80-88
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

import pointnet2.pointnet2 as ops
from ogc_b200 import backend as _backend_mod
from ogc_b200 import fp_fused, sa_fused

FORCE_COMPOSED = False   # tests: run the torch-composed SA path on the GPU to compare with the fused kernels
HEAD_SIDE_STREAM = True  # slot transformer + object MLP on a side stream underneath the feature-propagation chain
_HEAD_STREAMS = {}


def _head_stream(device):
    key = (device.type, device.index)
    if key not in _HEAD_STREAMS:
        _HEAD_STREAMS[key] = torch.cuda.Stream(device=device)
    return _HEAD_STREAMS[key]

GN_GROUPS = 4  # models/segnet_kitti.py:8  BN_CONFIG = GroupNorm, 4 groups

# (npoint divisor, radii, nsamples, mlps-without-xyz) per SA level; FP mlps listed in module order
# (FP_modules[0] is applied LAST, on the full-resolution cloud).  models/segnet_*.py:26-51
SEGNET_SPECS = {
    "kitti": {
        "sa": [(4, [1, 2], [64, 64], [[3, 32, 32, 32], [3, 32, 32, 64]]),
               (8, [4], [64], [[32 + 64, 64, 64, 128]]),
               (16, [8], [64], [[128, 128, 128, 256]])],
        "fp": [[64 + 3, 64, 64, 64], [32 + 64 + 128, 64, 64], [128 + 256, 128, 128]],
    },
    "sapien": {
        "sa": [(2, [0.1, 0.2], [64, 64], [[3, 64, 64, 64], [3, 64, 64, 128]]),
               (4, [0.4], [64], [[64 + 128, 128, 128, 256]])],
        "fp": [[128 + 3, 128, 128, 64], [256 + 64 + 128, 256, 128]],
    },
    "ogcdr": {
        "sa": [(2, [0.05, 0.1], [64, 64], [[3, 64, 64, 64], [3, 64, 64, 128]]),
               (4, [0.2], [64], [[64 + 128, 128, 128, 256]])],
        "fp": [[128 + 3, 128, 128, 64], [256 + 64 + 128, 256, 128]],
    },
}


def _pointwise_linear(weight, x, bias=None):
    """1x1 convolution as an fp32 matmul: weight (Cout,Cin,1[,1]), x (B,Cin,*) -> (B,Cout,*)."""
    w = weight.view(weight.shape[0], weight.shape[1])
    y = torch.matmul(w, x.flatten(2)).view(x.shape[0], w.shape[0], *x.shape[2:])
    if bias is not None:
        y = y + bias.view(1, -1, *([1] * (x.dim() - 2)))
    return y


class _Norm(nn.Module):
    """Holder named like the reference's utils/nn_util.py:6-11 wrapper: `<layer>.normlayer.gn.*`."""

    def __init__(self, channels):
        super().__init__()
        self.gn = nn.GroupNorm(GN_GROUPS, channels)


class ConvGNReLU(nn.Module):
    """conv1x1 (no bias when a norm follows, utils/nn_util.py:52) -> GroupNorm(4) -> ReLU."""

    def __init__(self, cin, cout, dim=2, norm=True, act=True):
        super().__init__()
        conv = nn.Conv2d if dim == 2 else nn.Conv1d
        self.conv = conv(cin, cout, kernel_size=1, bias=not norm)
        nn.init.kaiming_normal_(self.conv.weight)
        if not norm:
            nn.init.zeros_(self.conv.bias)
        self.normlayer = _Norm(cout) if norm else None
        self.act = act

    def forward(self, x):
        y = _pointwise_linear(self.conv.weight, x, self.conv.bias)
        if self.normlayer is not None:
            gn = self.normlayer.gn
            y = F.group_norm(y, GN_GROUPS, gn.weight, gn.bias, gn.eps)
        return F.relu(y) if self.act else y


class SharedMLP(nn.Module):
    """layer0 .. layerL-1 of ConvGNReLU, named as utils/nn_util.py:151-168."""

    def __init__(self, channels):
        super().__init__()
        self.n_layers = len(channels) - 1
        for i in range(self.n_layers):
            self.add_module(f"layer{i}", ConvGNReLU(channels[i], channels[i + 1]))

    def forward(self, x):
        for i in range(self.n_layers):
            x = getattr(self, f"layer{i}")(x)
        return x


class SetAbstraction(nn.Module):
    """One PointNet++ set-abstraction level (multi-scale grouping)."""

    def __init__(self, npoint, radii, nsamples, mlps):
        super().__init__()
        self.npoint, self.radii, self.nsamples = npoint, list(radii), list(nsamples)
        self.mlps = nn.ModuleList(SharedMLP([c[0] + 3] + c[1:]) for c in mlps)  # use_xyz adds 3 inputs

    def sample(self, xyz):
        """FPS + centre gather of this level: depends on the coordinates only (see MaskFormer3D.sample_chain)."""
        sel = ops.furthest_point_sample(xyz, self.npoint).long()
        return ops.gather_nd(xyz, sel).contiguous()

    def forward(self, xyz, features, new_xyz=None):
        """xyz (B,N,3), features (B,C,N) -> new_xyz (B,M,3), new_features (B,sum Cout,M).
        `new_xyz`: this level's centres when they were sampled ahead of time."""
        if new_xyz is None:
            new_xyz = self.sample(xyz)
        wait_ready(new_xyz)
        fused = (not FORCE_COMPOSED and xyz.is_cuda and getattr(_backend_mod.get_backend(), "name", "") == "b200"
                 and all(sa_fused.supported(ns, [m.layer0.conv.weight.shape[1]] +
                                            [getattr(m, f"layer{i}").conv.weight.shape[0] for i in range(m.n_layers)])
                         for ns, m in zip(self.nsamples, self.mlps)))
        if fused:
            feat_pm = features.transpose(1, 2).contiguous()          # point-major rows for coalesced gathers
        else:
            xyz_t = xyz.transpose(1, 2).contiguous()
            centre = new_xyz.transpose(1, 2).unsqueeze(-1)
        outs = []
        knn_cache = {}
        for radius, nsample, mlp in zip(self.radii, self.nsamples, self.mlps):
            if nsample not in knn_cache:   # identical k-NN for every scale with the same k
                same_k = [r for r, ns in zip(self.radii, self.nsamples) if ns == nsample]
                if fused and all(r is not None for r in same_k):
                    # every scale replaces neighbours beyond its radius by the nearest one: nothing farther than the
                    # largest radius can survive, so the search is bounded (centres are members of xyz)
                    knn_cache[nsample] = _backend_mod.get_backend().knn_bounded(nsample, new_xyz, xyz, max(same_k))
                else:
                    knn_cache[nsample] = ops.knn(nsample, new_xyz, xyz)
            dist, idx = knn_cache[nsample]
            idx = ops.clip_neighbours_by_radius(dist, idx, radius)
            if fused:
                layers = [(getattr(mlp, f"layer{i}").conv.weight, getattr(mlp, f"layer{i}").normlayer.gn.weight,
                           getattr(mlp, f"layer{i}").normlayer.gn.bias) for i in range(mlp.n_layers)]
                outs.append(sa_fused.fused_sa_mlp(xyz, new_xyz, feat_pm, idx, layers))
                continue
            grouped = torch.cat([ops.grouping_operation(xyz_t, idx) - centre,
                                 ops.grouping_operation(features, idx)], dim=1)   # (B,3+C,M,S)
            outs.append(mlp(grouped).max(dim=3).values)
        return new_xyz, torch.cat(outs, dim=1)


def wait_ready(t):
    """Tensors produced ahead of time on a side stream (train.SegTrainer._prefetch_geometry) carry the event they become
    ready at; the consumer's stream waits for it here (a no-op for ordinary tensors)."""
    ev = getattr(t, "_ogc_ready", None)
    if ev is not None:
        torch.cuda.current_stream().wait_event(ev)
    return t


class FeaturePropagation(nn.Module):
    def __init__(self, channels):
        super().__init__()
        self.mlp = SharedMLP(channels)

    def forward(self, unknown, known, unknown_feats, known_feats, nn=None):
        """unknown (B,n,3), known (B,m,3), unknown_feats (B,C1,n), known_feats (B,C2,m) -> (B,Cout,n).
        `nn`: (dist2, idx) of three_nn(unknown, known) when computed ahead of time (fused path only)."""
        widths = [self.mlp.layer0.conv.weight.shape[1]] + [getattr(self.mlp, f"layer{i}").conv.weight.shape[0]
                                                           for i in range(self.mlp.n_layers)]
        if (not FORCE_COMPOSED and unknown.is_cuda and fp_fused.supported(widths)
                and getattr(_backend_mod.get_backend(), "name", "") == "b200"):
            layers = [(getattr(self.mlp, f"layer{i}").conv.weight, getattr(self.mlp, f"layer{i}").normlayer.gn.weight,
                       getattr(self.mlp, f"layer{i}").normlayer.gn.bias) for i in range(self.mlp.n_layers)]
            if nn is not None:
                wait_ready(nn[0])
            return fp_fused.fused_fp(unknown, known, unknown_feats, known_feats, layers, nn)
        dist, idx = ops.three_nn(unknown.contiguous(), known.contiguous())
        recip = 1.0 / (dist + 1e-8)
        weight = recip / recip.sum(dim=2, keepdim=True)
        feats = ops.three_interpolate(known_feats.contiguous(), idx, weight.contiguous())
        if unknown_feats is not None:
            feats = torch.cat([feats, unknown_feats], dim=1)
        return self.mlp(feats.unsqueeze(-1)).squeeze(-1)


class _MaskHeadFn(torch.autograd.Function):
    """softmax_k(cos(point feature, slot) / temperature) on csrc/mask_head.cu; slots arrive normalised."""

    @staticmethod
    def forward(ctx, feats, slots_hat, inv_temp):
        import ctypes
        from ogc_b200 import _lib
        from ogc_b200.backend import TIMER
        be = _backend_mod.get_backend()
        feats, slots_hat = feats.contiguous(), slots_hat.contiguous()
        B, D, N = feats.shape
        K = slots_hat.shape[2]
        mask = torch.empty(B, N, K, dtype=torch.float32, device=feats.device)
        P = lambda t: ctypes.c_void_p(t.data_ptr())
        with TIMER.span("mask_head_fwd", B * N * 4 * (D + K)):
            _lib.check(be.lib.ogc_mask_head_fwd(B, D, N, K, float(inv_temp), P(feats), P(slots_hat), P(mask),
                                                ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "ogc_mask_head_fwd")
        be.launches += 1
        ctx.inv_temp = float(inv_temp)
        ctx.save_for_backward(feats, slots_hat, mask)
        return mask

    @staticmethod
    def backward(ctx, dmask):
        import ctypes
        from ogc_b200 import _lib
        from ogc_b200.backend import TIMER
        be = _backend_mod.get_backend()
        feats, slots_hat, mask = ctx.saved_tensors
        B, D, N = feats.shape
        K = slots_hat.shape[2]
        dmask = dmask.contiguous()
        dfeats = torch.empty_like(feats)
        dslots = torch.zeros_like(slots_hat)
        P = lambda t: ctypes.c_void_p(t.data_ptr())
        with TIMER.span("mask_head_bwd", B * N * 4 * (2 * D + 2 * K)):
            _lib.check(be.lib.ogc_mask_head_bwd(B, D, N, K, ctx.inv_temp, P(feats), P(slots_hat), P(mask), P(dmask),
                                                P(dfeats), P(dslots),
                                                ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "ogc_mask_head_bwd")
        be.launches += 1
        return dfeats, dslots, None


class _OutProj(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(dim, dim))
        self.bias = nn.Parameter(torch.zeros(dim))


class Attention(nn.Module):
    """Multi-head attention with nn.MultiheadAttention's parameter names (in_proj_weight,
    in_proj_bias, out_proj.weight, out_proj.bias) and arithmetic (batch_first, no dropout)."""

    def __init__(self, dim, heads):
        super().__init__()
        self.dim, self.heads = dim, heads
        self.in_proj_weight = nn.Parameter(torch.empty(3 * dim, dim))
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * dim))
        self.out_proj = _OutProj(dim)
        nn.init.xavier_uniform_(self.in_proj_weight)
        nn.init.kaiming_uniform_(self.out_proj.weight, a=math.sqrt(5))

    def forward(self, query, key, value):
        d, h = self.dim, self.heads
        wq, wk, wv = self.in_proj_weight.split(d, dim=0)
        bq, bk, bv = self.in_proj_bias.split(d, dim=0)
        B, Lq, _ = query.shape
        q = F.linear(query, wq, bq).view(B, Lq, h, d // h).transpose(1, 2)
        k = F.linear(key, wk, bk).view(B, -1, h, d // h).transpose(1, 2)
        v = F.linear(value, wv, bv).view(B, -1, h, d // h).transpose(1, 2)
        attn = torch.softmax(torch.matmul(q * (1.0 / math.sqrt(d // h)), k.transpose(-1, -2)), dim=-1)
        out = torch.matmul(attn, v).transpose(1, 2).reshape(B, Lq, d)
        return F.linear(out, self.out_proj.weight, self.out_proj.bias)


class SlotDecoderLayer(nn.Module):
    """Cross-attention, self-attention, feed-forward with pre-norm (utils/transformer_util.py:5-58)."""

    def __init__(self, dim, heads, hidden):
        super().__init__()
        self.norm_slot1 = nn.LayerNorm(dim)
        self.norm_slot2 = nn.LayerNorm(dim)
        self.norm_pre_ff = nn.LayerNorm(dim)
        self.cross_attn = Attention(dim, heads)
        self.self_attn = Attention(dim, heads)
        self.mlp = nn.Sequential(nn.Linear(dim, hidden), nn.ReLU(inplace=True), nn.Linear(hidden, dim))

    def forward(self, slot, feats):
        slot = slot + self.cross_attn(self.norm_slot1(slot), feats, feats)
        s = self.norm_slot2(slot)
        slot = slot + self.self_attn(s, s, s)
        return slot + self.mlp(self.norm_pre_ff(slot))


class MaskFormerHead(nn.Module):
    def __init__(self, n_slot, input_dim, n_layer, dim, heads, hidden):
        super().__init__()
        self.n_slot = n_slot
        self.query = nn.Embedding(n_slot, dim)
        self.mlp_input = nn.Sequential(nn.Linear(input_dim, dim), nn.ReLU(inplace=True), nn.Linear(dim, dim))
        self.norm_input = nn.LayerNorm(dim)
        self.transformer_layers = nn.ModuleList(SlotDecoderLayer(dim, heads, hidden) for _ in range(n_layer))

    def forward(self, point_feats):
        """point_feats (B,M,Cin) -> slots (B,K,D).  (The reference hard-codes .cuda() here,
        utils/transformer_util.py:110; we follow the input's device.)"""
        slot = self.query.weight.unsqueeze(0).expand(point_feats.shape[0], -1, -1)
        feats = self.norm_input(self.mlp_input(point_feats))
        for layer in self.transformer_layers:
            slot = layer(slot, feats)
        return slot


class MaskFormer3D(nn.Module):
    """pc (B,N,3), point_feats (B,N,3) -> soft object masks (B,N,K)."""

    def __init__(self, n_slot, n_point=8192, variant="kitti", n_transformer_layer=2, transformer_embed_dim=128,
                 use_xyz=True, transformer_input_pos_enc=False):
        super().__init__()
        if not use_xyz or transformer_input_pos_enc:
            raise NotImplementedError("only the configurations the reference ships (config/seg/*.yaml) are mirrored")
        spec = SEGNET_SPECS[variant]
        self.variant = variant
        self.n_slot = n_slot
        self.SA_modules = nn.ModuleList(
            SetAbstraction(int(n_point / div), radii, nsamples, [list(m) for m in mlps])
            for div, radii, nsamples, mlps in spec["sa"])
        self.FP_modules = nn.ModuleList(FeaturePropagation(list(c)) for c in spec["fp"])
        d = transformer_embed_dim
        self.MF_head = MaskFormerHead(n_slot, 256, n_transformer_layer, d, 8, d)
        self.object_mlp = nn.Sequential(ConvGNReLU(d, d, dim=1), ConvGNReLU(d, 64, dim=1, norm=False, act=False))

    def sample_chain(self, pc):
        """Centres of every SA level.  The FPS chain is a function of the coordinates alone and is latency-bound on
        B SMs (one CTA per cloud), so the trainer runs it on a side stream underneath wide kernels."""
        chain = []
        for sa in self.SA_modules:
            pc = sa.sample(pc)
            chain.append(pc)
        return chain

    def geometry_chain(self, pc):
        """sample_chain + the three_nn of every FP level: everything in the network that is a function of the
        coordinates alone.  -> (centres, fp_nn)"""
        centres = self.sample_chain(pc)
        l_pc = [pc] + centres
        be = _backend_mod.get_backend()
        fp_nn = [be.three_nn(l_pc[i].contiguous(), l_pc[i + 1].contiguous()) for i in range(len(self.FP_modules))]
        return centres, fp_nn

    def forward(self, pc, point_feats, centres=None, fp_nn=None):
        l_pc, l_feats = [pc], [point_feats.transpose(1, 2).contiguous()]
        for i, sa in enumerate(self.SA_modules):
            new_pc, new_feats = sa(l_pc[-1], l_feats[-1], None if centres is None else centres[i])
            l_pc.append(new_pc)
            l_feats.append(new_feats)
        fused_head = (not FORCE_COMPOSED and pc.is_cuda and self.n_slot <= 16
                      and getattr(_backend_mod.get_backend(), "name", "") == "b200")
        side = None
        if fused_head and HEAD_SIDE_STREAM:
            # The slot transformer + object MLP (~60 tiny kernels forward, ~150 backward, K = 10 slots) only need the
            # coarsest features, the feature-propagation chain only the set-abstraction outputs: two independent
            # branches.  The head runs on a side stream underneath the FP kernels; autograd replays every node on the
            # stream of its forward, so its backward overlaps the FP / SA backward the same way (and a stream capture
            # records the fork / join as parallel graph branches).
            main = torch.cuda.current_stream()
            side = _head_stream(pc.device)
            side.wait_stream(main)
            with torch.cuda.stream(side):
                slot = self.object_mlp(self.MF_head(l_feats[-1].transpose(1, 2)).transpose(1, 2))
                slot_hat = F.normalize(slot, dim=1)
        for i in range(len(self.FP_modules) - 1, -1, -1):     # coarse -> fine; FP_modules[i] lifts level i+1 to i
            l_feats[i] = self.FP_modules[i](l_pc[i], l_pc[i + 1], l_feats[i], l_feats[i + 1], None if fp_nn is None else fp_nn[i])
        if side is not None:
            torch.cuda.current_stream().wait_stream(side)
            if l_feats[0].shape[1] == 64:
                return _MaskHeadFn.apply(l_feats[0], slot_hat, 1.0 / 0.05)
        else:
            slot = self.MF_head(l_feats[-1].transpose(1, 2))                      # (B,K,D)
            slot = self.object_mlp(slot.transpose(1, 2))                          # (B,64,K)
        if fused_head and l_feats[0].shape[1] == 64:
            return _MaskHeadFn.apply(l_feats[0], F.normalize(slot, dim=1), 1.0 / 0.05)
        logits = torch.einsum("bdn,bdk->bnk", F.normalize(l_feats[0], dim=1), F.normalize(slot, dim=1)) / 0.05
        return logits.softmax(dim=-1)
