"""Fused BatchNorm shared MLP of the FlowStep3D blocks: autograd wrapper over csrc/bn_mlp.cu (+ the pointwise
contraction kernels of csrc/fp_mlp.cu and csrc/mlp_bwd.cu).

    fused_bn_mlp(grouped, convs, bns) -> (B, C_L, M)

computes the tail of the reference's PointNetSetAbstraction.forward / FlowEmbedding.forward
(utils/flowstep3d_util.py:126-137 and :52-64)

    max_s relu(BN(W_L ... relu(BN(W_1 grouped))))          (bns given: BatchNorm2d in TRAINING mode, batch statistics)
    max_s (W_1 grouped)                                    (bns None: the bare convolution blocks, use_act=False)

on a grouped tensor (B, C_in, M, S) with one contraction kernel per layer in the forward (normalisation + ReLU of the
previous layer folded into the operand loader) and two per layer in the backward; the per-channel statistics, the
tables the contraction kernels read, the pooling over nsample and the backward entry are the kernels of bn_mlp.cu.
No normalised / rectified tensor is materialised; only the pre-norm y_l are stored.  The BatchNorm running estimates
and `num_batches_tracked` are updated as nn.BatchNorm2d does.
"""
import ctypes
import os

import torch
from torch.autograd import Function

from . import _lib
from .backend import TIMER, get_backend


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _st():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def supported(widths, nsample):
    """Shapes the kernels cover: layer outputs multiples of 16 up to 256, fewer than 255 neighbour slots."""
    return nsample < 255 and all(c % 16 == 0 and c <= 256 for c in widths)


USE_TMA = {"0": False, "1": True}.get(os.environ.get("OGC_BN_TMA", "bwd"), "bwd")
# The dense inner layers (l >= 1) can run the TMA-staged tensor-core kernels of the SA block (csrc/sa_*_tma.cu, 3xTF32
# split; the M*S positions of a sample are (M*S)/64 "centres" of 64 positions to them): False = nowhere, "bwd" = weight /
# input gradients only, True = forward too.  Measured on a B200 with True: the step 27.3 -> 23.7 ms, every block within
# 5e-7 of fp64 either way -- but the untrained recurrent network amplifies the split's different rounding in the FORWARD
# to 1.3e-4 on the second iteration's flow (fp32 SIMT kernels: 5e-5; parity bound 1e-4).  "bwd" (the default) leaves every
# forward value and decision (ReLU masks, arg-max slots, neighbour sets) untouched -- same golden errors as False, flows and
# gradients -- and keeps most of the gain: 24.0 ms.


def _tma_ok(P, cin, cout, fwd=False):
    from . import sa_fused
    if not (USE_TMA is True or (USE_TMA == "bwd" and not fwd)):
        return False
    return sa_fused.USE_TC and P % 128 == 0 and cin % 32 == 0 and cin <= 128 and cout % 32 == 0 and cout <= 128


class _FusedBnMlp(Function):
    @staticmethod
    def forward(ctx, x, use_act, running, no_grad_rows, *params):
        """x (B,Cin,M,S); params = (W_1, gamma_1, beta_1, ...) with use_act, (W_1, ..., W_L) without;
        running = [(running_mean, running_var, momentum) or None per layer] (updated in place); the first no_grad_rows
        channels of x get a zero gradient (the caller knows nothing upstream of them needs one)."""
        be = get_backend()
        lib = be.lib
        stride = 3 if use_act else 1
        L = len(params) // stride
        B, cin0, M, S = x.shape
        P = M * S
        dev = x.device
        f32 = dict(dtype=torch.float32, device=dev)
        cmax = max(params[stride * l].shape[0] for l in range(L))
        sums_all = torch.zeros(L, cmax, 2, dtype=torch.float64, device=dev)        # one fill for all layers
        gn_scratch = torch.empty(B, 4, 2, dtype=torch.float64, device=dev)         # the contraction kernel's GroupNorm sums: unused
        ys, sss, mrs = [], [], []
        a_prev, ss_prev = x, None
        for l in range(L):
            W = params[stride * l]
            cout, cin = W.shape[0], W.shape[1]
            y = torch.empty(B, cout, P, **f32)
            tma = l > 0 and use_act and _tma_ok(P, cin, cout, fwd=True)
            with TIMER.span("flow_mlp_fwd_tma" if tma else "flow_mlp_fwd", B * 4 * P * (cin + cout), 2 * B * P * cin * cout):
                if tma:
                    w2d = W.detach().reshape(cout, cin).contiguous()
                    _lib.check(lib.ogc_sa_fwd_tma(B, P // 64, 64, cin, cout, 0, _p(a_prev), _p(ss_prev), _p(w2d), _p(y), _p(gn_scratch),
                                                  None, None, None, None, _st()), "ogc_sa_fwd_tma")
                else:
                    wt = W.detach().reshape(cout, cin).t().contiguous()       # only the SIMT kernel wants W^T
                    _lib.check(lib.ogc_pw_mlp_layer_fwd(B, P, cin, cout, _p(a_prev), _p(ss_prev), _p(wt), _p(y), _p(gn_scratch), _st()),
                               "ogc_pw_mlp_layer_fwd")
            be.launches += 1
            ys.append(y)
            if use_act:
                gamma, beta = params[3 * l + 1], params[3 * l + 2]
                ss = torch.empty(B, cout, 2, **f32)
                mr = torch.empty(cout, 2, **f32)
                rm, rv, mom = running[l] if running[l] is not None else (None, None, 0.0)
                with TIMER.span("flow_bn_stats", B * 4 * P * cout):
                    _lib.check(lib.ogc_bn_stats(B, cout, P, _p(y), _p(sums_all[l]), _st()), "ogc_bn_stats")
                _lib.check(lib.ogc_bn_finalize(B, cout, B * P, _p(sums_all[l]), _p(gamma.detach()), _p(beta.detach()), _p(ss),
                                               _p(mr), _p(rm), _p(rv), float(mom), _st()), "ogc_bn_finalize")
                be.launches += 2
                sss.append(ss); mrs.append(mr)
                a_prev, ss_prev = y, ss
            else:
                a_prev, ss_prev = y, None
        cL = ys[-1].shape[1]
        out = torch.empty(B, cL, M, **f32)
        sel = torch.empty(B, cL, M, dtype=torch.uint8, device=dev)
        with TIMER.span("flow_pool", B * cL * (4 * P + 5 * M)):
            _lib.check(lib.ogc_bn_pool(B, cL, M, S, _p(ys[-1]), _p(sss[-1]) if use_act else None, _p(out), _p(sel), _st()),
                       "ogc_bn_pool")
        be.launches += 1
        ctx.dims = (B, M, S, L, use_act, no_grad_rows)
        ctx.param_objs = params
        ctx.save_for_backward(x, sel, *ys, *sss, *mrs, *[p.detach() for p in params])
        return out

    @staticmethod
    def backward(ctx, go):
        be = get_backend()
        lib = be.lib
        B, M, S, L, use_act, skip = ctx.dims
        P = M * S
        stride = 3 if use_act else 1
        saved = ctx.saved_tensors
        x, sel = saved[:2]
        ys = saved[2:2 + L]
        nl = L if use_act else 0
        sss, mrs = saved[2 + L:2 + L + nl], saved[2 + L + nl:2 + L + 2 * nl]
        params = saved[2 + L + 2 * nl:]
        dev = x.device
        f32 = dict(dtype=torch.float32, device=dev)
        go = go.contiguous()
        from . import sa_fused
        tg = sa_fused.grad_targets(ctx.param_objs)      # accumulate straight into the parameters' .grad (trainer's backward)
        grads = [None] * len(params)
        cL = ys[-1].shape[1]
        cmax = max(y.shape[1] for y in ys)
        dz = torch.empty(B, cL, P, **f32)
        ab_all = torch.zeros(L, cmax, 2, dtype=torch.float64, device=dev) if use_act else None
        with TIMER.span("flow_pool_bwd", B * cL * (4 * P + 9 * M)):
            _lib.check(lib.ogc_bn_pool_bwd(B, cL, M, S, _p(go), _p(sel), _p(ys[-1]), _p(mrs[-1]) if use_act else None, _p(dz),
                                           _p(ab_all[L - 1]) if use_act else None, _st()), "ogc_bn_pool_bwd")
        be.launches += 1
        if use_act:       # the dense input-gradient kernel also emits GroupNorm sums of the previous layer: parked here
            gn_mr = torch.zeros(B, 4, 2, **f32)
            gn_ab = torch.empty(B, 4, 2, dtype=torch.float64, device=dev)
            gn_dg = torch.empty(2, cmax, **f32)
        d_x = None
        for l in range(L - 1, -1, -1):
            W = params[stride * l]
            cout, cin = W.shape[0], W.shape[1]
            w2d = W.reshape(cout, cin).contiguous()
            coef = torch.empty(B, cout, 4, **f32)
            if use_act:
                gamma = params[3 * l + 1]
                dgamma = tg[3 * l + 1] if tg else torch.zeros(cout, **f32)
                dbeta = tg[3 * l + 2] if tg else torch.zeros(cout, **f32)
                _lib.check(lib.ogc_bn_bwd_coef(B, cout, B * P, _p(ab_all[l]), _p(mrs[l]), _p(gamma), _p(coef), _p(dgamma),
                                               _p(dbeta), _st()), "ogc_bn_bwd_coef")
                be.launches += 1
                if not tg:
                    grads[3 * l + 1], grads[3 * l + 2] = dgamma, dbeta
            else:
                coef.zero_()
                coef[:, :, 0].fill_(1.0)                                                    # dY = dz
            dW = tg[stride * l].view(cout, cin) if tg else torch.zeros(cout, cin, **f32)
            a_prev = ys[l - 1] if l else x
            ss_prev = sss[l - 1] if (l and use_act) else None
            tma = l > 0 and use_act and _tma_ok(P, cin, cout)
            with TIMER.span("flow_mlp_dw_tma" if tma else "flow_mlp_dw", B * P * 4 * (2 * cout + cin), 2 * B * P * cin * cout):
                if tma:
                    _lib.check(lib.ogc_sa_dw_tma(B, P // 64, 64, cout, cin, _p(dz), None, 0, 0, None, _p(ys[l]), _p(coef),
                                                 _p(a_prev), _p(ss_prev), _p(dW), _st()), "ogc_sa_dw_tma")
                else:
                    _lib.check(lib.ogc_sa_mlp_layer_dw(B, 0, P, 1, cout, cin, 0, _p(dz), None, 0, 0, None, _p(ys[l]), _p(coef),
                                                       _p(a_prev), _p(ss_prev), None, None, None, None, _p(dW), _st()),
                               "ogc_sa_mlp_layer_dw")
            be.launches += 1
            if not tg:
                grads[stride * l] = dW.view_as(W)
            if l > 0 and use_act:
                cprev = ys[l - 1].shape[1]
                dz_prev = torch.empty(B, cprev, P, **f32)
                tma = _tma_ok(P, cprev, cout)
                with TIMER.span("flow_mlp_dx_tma" if tma else "flow_mlp_dx", B * P * 4 * (2 * cout + 2 * cprev), 2 * B * P * cprev * cout):
                    if tma:
                        _lib.check(lib.ogc_sa_dx_tma(
                            B, P // 64, 64, cout, cin, 0, cprev, _p(dz), None, 0, 0, None, _p(ys[l]), _p(coef), _p(w2d),
                            _p(ys[l - 1]), _p(sss[l - 1]), _p(gn_mr), _p(params[3 * (l - 1) + 1]), _p(dz_prev),
                            _p(gn_ab), _p(gn_dg[0]), _p(gn_dg[1]), _st()), "ogc_sa_dx_tma")
                    else:
                        _lib.check(lib.ogc_sa_mlp_layer_dx(
                            B, 0, P, 1, cout, cin, 0, cprev, _p(dz), None, 0, 0, None, _p(ys[l]), _p(coef), _p(w2d),
                            _p(ys[l - 1]), _p(sss[l - 1]), _p(gn_mr), _p(params[3 * (l - 1) + 1]), _p(dz_prev),
                            _p(gn_ab), _p(gn_dg[0]), _p(gn_dg[1]), None, None, 0, 0, _st()), "ogc_sa_mlp_layer_dx")
                with TIMER.span("flow_bn_bwd_stats", B * P * 8 * cprev):
                    _lib.check(lib.ogc_bn_bwd_stats(B, cprev, P, _p(dz_prev), _p(ys[l - 1]), _p(mrs[l - 1]), _p(ab_all[l - 1]),
                                                    _st()), "ogc_bn_bwd_stats")
                be.launches += 2
                dz = dz_prev
            elif l > 0:
                # bare convolutions stacked without activation: dz_{l-1} = W_l^T dz_l
                cprev = ys[l - 1].shape[1]
                dz_prev = torch.empty(B, cprev, P, **f32)
                for off in range(0, cprev, 128):
                    rows = min(128, cprev - off)
                    _lib.check(lib.ogc_pw_mlp_input_grad(B, P, cout, cin, off, rows, _p(dz), _p(ys[l]), _p(coef), _p(w2d),
                                                         _p(dz_prev), cprev, off, _st()), "ogc_pw_mlp_input_grad")
                    be.launches += 1
                dz = dz_prev
            elif ctx.needs_input_grad[0]:
                # rows [skip, cin) of W^T dY; 3 + 64 input channels would otherwise land in a 128-row tile (profiles/README.md)
                d_x = torch.empty(B, cin, P, **f32)
                if skip:
                    d_x[:, :skip].zero_()
                for off in range(skip, cin, 128):
                    rows = min(128, cin - off)
                    with TIMER.span("flow_mlp_dx", B * P * 4 * (2 * cout + rows), 2 * B * P * rows * cout):
                        _lib.check(lib.ogc_pw_mlp_input_grad(B, P, cout, cin, off, rows, _p(dz), _p(ys[0]), _p(coef), _p(w2d),
                                                             _p(d_x), cin, off, _st()), "ogc_pw_mlp_input_grad")
                    be.launches += 1
                d_x = d_x.view(B, cin, M, S)
        return (d_x, None, None, None, *grads)


def fused_bn_mlp(grouped, convs, bns, no_grad_rows=0):
    """grouped (B,Cin,M,S); convs = the block's nn.Conv2d 1x1 (bias=False) modules; bns = its nn.BatchNorm2d modules
    (training mode: batch statistics, running estimates updated) or None for a bare convolution block.
    no_grad_rows: the leading channels of `grouped` whose gradient nobody needs (the centred coordinates of a cloud that
    does not require grad): their gradient is returned as zeros instead of being computed."""
    use_act = bns is not None
    params, running = [], []
    for l, conv in enumerate(convs):
        params.append(conv.weight)
        if use_act:
            bn = bns[l]
            params += [bn.weight, bn.bias]
            if bn.track_running_stats and bn.running_mean is not None:
                if bn.momentum is None:
                    raise NotImplementedError("cumulative-average BatchNorm (momentum=None) is not used by the reference")
                running.append((bn.running_mean, bn.running_var, bn.momentum))
                bn.num_batches_tracked.add_(1)
            else:
                running.append(None)
    return _FusedBnMlp.apply(grouped.contiguous(), use_act, running, int(no_grad_rows), *params)
