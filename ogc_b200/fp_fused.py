"""Fused feature-propagation block: autograd wrapper over csrc/fp_mlp.cu (+ the dense-mode kernels of
csrc/mlp_bwd.cu for the layer backward).

    fused_fp(unknown, known, skip, known_feats, layers) -> (B, C_L, n)

computes the reference's PointnetFPModule.forward (utils/pointnet2_util.py:96-120)

    relu(GN(W_L ... relu(GN(W_1 [three_interpolate(known_feats, idx, w) ; skip]))))

with one kernel for the inverse-distance weights + interpolation + concat, one kernel per layer in the forward
(GroupNorm+ReLU of the previous layer folded into the operand loader, statistics in the epilogue) and two per layer
in the backward.  No normalised / rectified intermediate is materialised; only the pre-norm y_l are stored.
"""
import ctypes

import torch
from torch.autograd import Function

from . import _lib
from .backend import TIMER, get_backend


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _st():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def supported(channels):
    """Layer widths the kernels cover: outputs multiples of 16 up to 256 (GroupNorm(4) rows per tile)."""
    return all(c % 16 == 0 and c <= 256 for c in channels[1:])


USE_TMA = True      # dense layers (l >= 1) through the TMA-staged tensor-core kernels of the SA block (csrc/sa_*_tma.cu): the
                    # n points of a cloud are n / 64 "centres" of 64 positions to them


def _tma_ok(n, cin, cout):
    from . import sa_fused
    return (USE_TMA and sa_fused.USE_TC and n % 128 == 0 and cin % 32 == 0 and cin <= 128 and cout % 32 == 0 and cout <= 128)


class _FusedFP(Function):
    @staticmethod
    def forward(ctx, unknown, known, skip, known_feats, nn_d2, nn_idx, *params):
        """params = (W_1, gamma_1, beta_1, ..., W_L, gamma_L, beta_L); W_l (Cout,Cin,1,1)."""
        be = get_backend()
        lib = be.lib
        L = len(params) // 3
        B, n, _ = unknown.shape
        m = known.shape[1]
        c2 = known_feats.shape[1]
        c1 = 0 if skip is None else skip.shape[1]
        dev = unknown.device
        f32 = dict(dtype=torch.float32, device=dev)
        d2, idx = (nn_d2, nn_idx) if nn_idx is not None else be.three_nn(unknown, known)
        x = torch.empty(B, c2 + c1, n, **f32)
        wgt = torch.empty(B, n, 3, **f32)
        with TIMER.span("fp_interp_concat", B * (4 * c2 * m + 24 * n + 4 * c1 * n + 4 * (c2 + c1) * n + 12 * n)):
            _lib.check(lib.ogc_fp_interp_concat(B, c2, m, c1, n, _p(known_feats), _p(idx), _p(d2), _p(skip), _p(x),
                                                _p(wgt), _st()), "ogc_fp_interp_concat")
        be.launches += 1
        ys, sss, mrs = [], [], []
        a_prev, ss_prev = x, None
        sums_all = torch.zeros(L, B, 4, 2, dtype=torch.float64, device=dev)      # one fill for all layers
        for l in range(L):
            W, gamma, beta = params[3 * l], params[3 * l + 1], params[3 * l + 2]
            cout, cin = W.shape[0], W.shape[1]
            y = torch.empty(B, cout, n, **f32)
            sums = sums_all[l]
            with TIMER.span(f"fp_mlp_fwd[{cin}>{cout}]" if TIMER.detail else "fp_mlp_fwd", B * 4 * n * (cin + cout)):
                if l > 0 and _tma_ok(n, cin, cout):
                    w2d = W.detach().reshape(cout, cin).contiguous()
                    _lib.check(lib.ogc_sa_fwd_tma(B, n // 64, 64, cin, cout, 0, _p(a_prev), _p(ss_prev), _p(w2d), _p(y), _p(sums),
                                                  None, None, None, None, _st()), "ogc_sa_fwd_tma")
                else:
                    wt = W.detach().reshape(cout, cin).t().contiguous()       # only the SIMT kernel wants W^T
                    _lib.check(lib.ogc_pw_mlp_layer_fwd(B, n, cin, cout, _p(a_prev), _p(ss_prev), _p(wt), _p(y), _p(sums),
                                                        _st()), "ogc_pw_mlp_layer_fwd")
            ss = torch.empty(B, cout, 2, **f32)
            mr = torch.empty(B, 4, 2, **f32)
            _lib.check(lib.ogc_gn_finalize(B, cout, (cout // 4) * n, _p(sums), _p(gamma.detach()), _p(beta.detach()),
                                           _p(ss), _p(mr), _st()), "ogc_gn_finalize")
            be.launches += 2
            ys.append(y); sss.append(ss); mrs.append(mr)
            a_prev, ss_prev = y, ss
        cL = ys[-1].shape[1]
        out = torch.empty(B, cL, n, **f32)
        _lib.check(lib.ogc_gn_relu_apply(B, cL, n, _p(ys[-1]), _p(sss[-1]), _p(out), _st()), "ogc_gn_relu_apply")
        be.launches += 1
        ctx.dims = (B, n, m, c2, c1, L)
        ctx.param_objs = params
        ctx.save_for_backward(idx, wgt, x, *ys, *sss, *mrs, *[p.detach() for p in params])
        return out

    @staticmethod
    def backward(ctx, go):
        be = get_backend()
        lib = be.lib
        B, n, m, c2, c1, L = ctx.dims
        saved = ctx.saved_tensors
        idx, wgt, x = saved[:3]
        ys, sss, mrs = saved[3:3 + L], saved[3 + L:3 + 2 * L], saved[3 + 2 * L:3 + 3 * L]
        params = saved[3 + 3 * L:]
        dev = x.device
        f32 = dict(dtype=torch.float32, device=dev)
        go = go.contiguous()
        grads = [None] * (3 * L)
        cL = ys[-1].shape[1]
        dz = torch.empty(B, cL, n, **f32)
        from . import sa_fused
        tg = sa_fused.grad_targets(ctx.param_objs)          # accumulate straight into the parameters' .grad (trainer's backward)
        ab_all = torch.zeros(L, B, 4, 2, dtype=torch.float64, device=dev)        # one fill for all layers
        ab = ab_all[L - 1]
        dgamma = tg[3 * (L - 1) + 1] if tg else torch.zeros(cL, **f32)
        dbeta = tg[3 * (L - 1) + 2] if tg else torch.zeros(cL, **f32)
        _lib.check(lib.ogc_gn_relu_bwd_stats(B, cL, n, _p(go), _p(ys[-1]), _p(sss[-1]), _p(mrs[-1]),
                                             _p(params[3 * (L - 1) + 1]), _p(dz), _p(ab), _p(dgamma), _p(dbeta), _st()),
                   "ogc_gn_relu_bwd_stats")
        be.launches += 1
        d_known = d_skip = None
        for l in range(L - 1, -1, -1):
            W, gamma = params[3 * l], params[3 * l + 1]
            cout, cin = W.shape[0], W.shape[1]
            w2d = W.reshape(cout, cin).contiguous()
            coef = torch.empty(B, cout, 4, **f32)
            _lib.check(lib.ogc_gn_bwd_coef(B, cout, (cout // 4) * n, _p(ab), _p(mrs[l]), _p(gamma), _p(coef), _st()),
                       "ogc_gn_bwd_coef")
            if not tg:
                grads[3 * l + 1], grads[3 * l + 2] = dgamma, dbeta
            dW = tg[3 * l].view(cout, cin) if tg else torch.zeros(cout, cin, **f32)
            a_prev = ys[l - 1] if l else x
            ss_prev = sss[l - 1] if l else None
            with TIMER.span(f"fp_mlp_dw[{cin}>{cout}]" if TIMER.detail else "fp_mlp_dw", B * n * 4 * (2 * cout + cin)):
                if l > 0 and _tma_ok(n, cin, cout):
                    _lib.check(lib.ogc_sa_dw_tma(B, n // 64, 64, cout, cin, _p(dz), None, 0, 0, None, _p(ys[l]), _p(coef),
                                                 _p(a_prev), _p(ss_prev), _p(dW), _st()), "ogc_sa_dw_tma")
                else:
                    _lib.check(lib.ogc_sa_mlp_layer_dw(B, 0, n, 1, cout, cin, 0, _p(dz), None, 0, 0, None, _p(ys[l]), _p(coef),
                                                       _p(a_prev), _p(ss_prev), None, None, None, None, _p(dW), _st()),
                               "ogc_sa_mlp_layer_dw")
            be.launches += 2
            if not tg:
                grads[3 * l] = dW.view_as(W)
            if l > 0:
                cprev = params[3 * (l - 1)].shape[0]
                dz_prev = torch.empty(B, cprev, n, **f32)
                ab_prev = ab_all[l - 1]
                dgamma_prev = tg[3 * (l - 1) + 1] if tg else torch.zeros(cprev, **f32)
                dbeta_prev = tg[3 * (l - 1) + 2] if tg else torch.zeros(cprev, **f32)
                with TIMER.span(f"fp_mlp_dx[{cout}>{cprev}]" if TIMER.detail else "fp_mlp_dx", B * n * 4 * (2 * cout + 2 * cprev)):
                    if _tma_ok(n, cprev, cout):
                        _lib.check(lib.ogc_sa_dx_tma(
                            B, n // 64, 64, cout, cin, 0, cprev, _p(dz), None, 0, 0, None, _p(ys[l]), _p(coef), _p(w2d),
                            _p(ys[l - 1]), _p(sss[l - 1]), _p(mrs[l - 1]), _p(params[3 * (l - 1) + 1]), _p(dz_prev),
                            _p(ab_prev), _p(dgamma_prev), _p(dbeta_prev), _st()), "ogc_sa_dx_tma")
                    else:
                        _lib.check(lib.ogc_sa_mlp_layer_dx(
                            B, 0, n, 1, cout, cin, 0, cprev, _p(dz), None, 0, 0, None, _p(ys[l]), _p(coef), _p(w2d),
                            _p(ys[l - 1]), _p(sss[l - 1]), _p(mrs[l - 1]), _p(params[3 * (l - 1) + 1]), _p(dz_prev),
                            _p(ab_prev), _p(dgamma_prev), _p(dbeta_prev), None, None, 0, 0, _st()), "ogc_sa_mlp_layer_dx")
                be.launches += 1
                dz, ab, dgamma, dbeta = dz_prev, ab_prev, dgamma_prev, dbeta_prev
            else:
                parts = []
                if ctx.needs_input_grad[3]:
                    parts.append((0, c2))
                if c1 > 0 and ctx.needs_input_grad[2]:
                    parts.append((c2, c1))
                for start, width in parts:
                    dx = torch.empty(B, width, n, **f32)
                    for off in range(0, width, 128):
                        rows = min(128, width - off)
                        with TIMER.span(f"fp_mlp_dx[{cout}>in{rows}]" if TIMER.detail else "fp_mlp_dx", B * n * 4 * (2 * cout + rows)):
                            _lib.check(lib.ogc_pw_mlp_input_grad(B, n, cout, cin, start + off, rows, _p(dz), _p(ys[0]),
                                                                 _p(coef), _p(w2d), _p(dx), width, off, _st()),
                                       "ogc_pw_mlp_input_grad")
                        be.launches += 1
                    if start == 0:
                        d_known = be.three_interpolate_grad(dx, idx, wgt, m)
                    else:
                        d_skip = dx
        return (None, None, d_skip, d_known, None, None, *grads)


def fused_fp(unknown, known, skip, known_feats, layers, nn=None):
    """unknown (B,n,3), known (B,m,3), skip (B,C1,n) or None, known_feats (B,C2,m),
    layers = [(W, gamma, beta), ...]  ->  (B, C_L, n).  nn = (dist2, idx) of three_nn(unknown, known) when it was
    computed ahead of time."""
    flat = [t for layer in layers for t in layer]
    d2, idx = nn if nn is not None else (None, None)
    return _FusedFP.apply(unknown.contiguous(), known.contiguous(), None if skip is None else skip.contiguous(),
                          known_feats.contiguous(), d2, idx, *flat)
