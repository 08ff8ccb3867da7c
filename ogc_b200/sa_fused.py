"""Fused set-abstraction MLP: autograd wrapper over csrc/mlp.cu + csrc/mlp_bwd.cu.

    fused_sa_mlp(xyz, new_xyz, feat_pm, idx, layers) -> (B, C_L, M)

computes, for one grouper of a PointNet++ SA level (utils/pointnet2_util.py:33-44),
    max_s  relu(GN(W_L ... relu(GN(W_1 [xyz_j - centre ; feat_j])))),   j = idx[b, m, s]
without materialising the grouped (B,3+C,M,S) tensor or any normalised / rectified activation:
one kernel per layer in the forward, two per layer in the backward (see the .cu headers).
"""
import ctypes
import os

import torch
from torch.autograd import Function

from . import _lib
from .backend import TIMER, get_backend


USE_TC = True   # tcgen05 3xTF32 kernels for the layers they cover (tests flip it to compare with the SIMT kernels)


def _tc_ok(S, cin, cout, gather):
    k = cin - 3 if gather else cin
    return USE_TC and S == 64 and cout >= 64 and cout % 16 == 0 and cout <= 256 and 32 <= k <= 128 and k % 4 == 0


USE_NARROW = True   # warp-per-centre kernels (csrc/mlp_narrow.cu) for the dense 32 -> 32 / 32 -> 64 layers of SA level 1


def _narrow_ok(S, cin, cout):
    return USE_NARROW and S == 64 and cin == 32 and cout in (32, 64)


TC_DW_ALL = False   # tests: route every supported dW through the tensor-core kernel


def _tc_dw_ok(S, cin, cout):
    # measured (scratch/mlp_bench.py, dY operand through tensor memory + cp.async ring): the tensor-core dW wins from
    # 64 input channels up (64 -> 64: 0.39 vs 0.56 ms SIMT, 99 -> 64: 0.58 vs 1.12, 128 -> 256: 0.71 vs 1.61); for the
    # 32-channel layers of SA level 1 it is on par with or behind the SIMT split-K kernel (0.67-0.90 vs 0.62-0.93)
    if TC_DW_ALL:
        return USE_TC and S == 64 and 32 <= cin <= 160 and 32 <= cout <= 256 and (cin % 4 == 0 or (cin - 3) % 4 == 0)
    return USE_TC and S == 64 and 64 <= cin <= 160 and 32 <= cout <= 256 and (cin % 4 == 0 or (cin - 3) % 4 == 0)


def _tc_dx_ok(S, cout, rows, scatter):
    return USE_TC and S == 64 and 32 <= cout <= (128 if scatter else 256) and rows <= 128 and (scatter or rows % 16 == 0)


USE_CHAIN_DX = True   # input-gradient kernel in the round-2 orientation (csrc/sa_chain_bwd.cu: positions on the MMA's M axis,
                      # operands staged by tensor-map TMA): 1.2-2.2x the per-layer kernel at KITTI-SF sizes


def _chain_dx_ok(S, M, cout, rows, scatter):
    if USE_CHAIN_DX in ("dense", "scatter") and USE_CHAIN_DX != ("scatter" if scatter else "dense"):
        return False        # diagnostics: only one of the two modes
    return (USE_TC and bool(USE_CHAIN_DX) and S == 64 and M % 2 == 0 and cout % 32 == 0 and cout <= 256 and rows <= 128
            and rows % (16 if scatter else 32) == 0 and 8 * rows * cout <= 180 * 1024)   # resident W^T (hi + lo) fits one SM


USE_DW_TMA = True     # dense-layer weight gradient with TMA-staged operands (csrc/sa_dw_tma.cu)
DW_TMA_NARROW = True  # ... also for SA level 1's 32 -> 32 layers (0.22-0.25 ms vs 0.22-0.28 for the warp-per-centre kernel)


def _dw_tma_ok(S, M, cin, cout):
    return (USE_TC and USE_DW_TMA and S == 64 and M % 2 == 0 and cout % 32 == 0 and cout <= 256 and cin % 32 == 0 and cin <= 128
            and (DW_TMA_NARROW or cin >= 64 or cout >= 64))


USE_FWD_TMA = True    # dense-layer forward with the TMA-staged operand (csrc/sa_fwd_tma.cu)
FWD_TMA_NARROW = False # ... also for SA level 1's 32 -> 32 layers (measured: the warp-per-centre kernel is faster there)


def _fwd_tma_ok(S, M, cin, cout):
    return (USE_TC and USE_FWD_TMA and S == 64 and M % 2 == 0 and cout % 32 == 0 and cout <= 256 and cin % 32 == 0 and cin <= 128
            and (FWD_TMA_NARROW or cin >= 64 or cout >= 64))


USE_DX_TMA = "auto"   # dense input gradient in the channel-major TMA style (csrc/sa_dx_tma.cu): True = every dense layer,
                      # "synth" = only the last layer (dz synthesised from the pooled gradient), "auto" = the last layer and the
                      # 128-wide dense layers (measured: 64 -> 64 is faster in the positions-on-M kernel), False = never


def _dx_tma_ok(S, cout, rows, synth):
    if not (USE_TC and USE_DX_TMA) or (USE_DX_TMA == "synth" and not synth) or (USE_DX_TMA == "auto" and not synth and cout < 128):
        return False
    return S == 64 and rows % 32 == 0 and rows <= 128 and cout % 32 == 0 and (cout <= 128 or cout == 256)


ACCUMULATE_INTO_GRAD = False   # set by the trainers around loss.backward(): the fused blocks' weight / GroupNorm gradients are
                               # accumulated by the kernels straight into the parameters' .grad (views of the flat gradient
                               # buffer) instead of into fresh zeroed tensors that autograd then adds on: ~115 tiny launches less


def grad_targets(param_objs):
    """The .grad tensors of the parameters if every one can take the kernels' atomic accumulation directly, else None."""
    if not ACCUMULATE_INTO_GRAD:
        return None
    gs = [p.grad for p in param_objs]
    if any(g is None or not g.is_contiguous() or g.dtype != torch.float32 or not p.requires_grad for g, p in zip(gs, param_objs)):
        return None
    return gs


DEBUG_KEEP = None     # diagnostics: a list that collects (layer, dz_prev, ab_prev, coef) of every backward
STORE_Y = True      # stage 1: keep the pre-norm tensors for the per-layer backward kernels
USE_CHAIN = False   # round-2 kernels (csrc/sa_chain_*.cu): positions on the MMA's M axis, layers chained through TMEM.
                    # Correct (tests/test_gpu_sa_chain.py) but not yet faster than the per-layer kernels at KITTI-SF sizes
                    # (profiles/README.md, round 2): kept behind this switch


def _chain_plan(M, S, Cf, widths):
    """How the chained forward covers a grouper: "full" = all layers resident (L passes, nothing stored),
    "layer" = one layer per launch with the pre-norm tensors stored once, None = not covered."""
    if not USE_CHAIN or S != 64 or M % 2:
        return None
    lib = get_backend().lib
    arr = (ctypes.c_int * 3)(*(list(widths) + [0, 0])[:3])
    if lib.ogc_sa_chain_fits(M, S, Cf, 1, len(widths), arr):
        return "full"
    if not lib.ogc_sa_chain_fits(M, S, Cf, 1, 1, arr):
        return None
    for l in range(1, len(widths)):
        one = (ctypes.c_int * 3)(widths[l], 0, 0)
        if not lib.ogc_sa_chain_fits(M, S, widths[l - 1], 0, 1, one):
            return None
    return "layer"


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _st():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


class _FusedSAMLP(Function):
    @staticmethod
    def forward(ctx, xyz, new_xyz, feat_pm, idx, *params):
        """params = (W_1, gamma_1, beta_1, ..., W_L, gamma_L, beta_L); W_l (Cout,Cin,1,1)."""
        be = get_backend()
        lib = be.lib
        L = len(params) // 3
        B, N, _ = xyz.shape
        M, S = idx.shape[1], idx.shape[2]
        P = M * S
        dev = xyz.device
        Cf = 0 if feat_pm is None else feat_pm.shape[2]
        f32 = dict(dtype=torch.float32, device=dev)
        ys, sss, mrs = [], [], []
        y_prev = ss_prev = None
        widths = [params[3 * l].shape[0] for l in range(L)]
        plan = _chain_plan(M, S, Cf, widths) if feat_pm is not None and L <= 3 else None
        if plan is not None:
            cL = widths[-1]
            w2d = [params[3 * l].detach().reshape(widths[l], -1).contiguous() for l in range(L)]
            hx = torch.empty(B, 2 * M, cL, **f32)
            hn = torch.empty(B, 2 * M, cL, **f32)
            ax = torch.empty(B, 2 * M, cL, dtype=torch.uint8, device=dev)
            an = torch.empty(B, 2 * M, cL, dtype=torch.uint8, device=dev)
            for l in range(L):
                last = l == L - 1
                y = torch.empty(B, widths[l], P, **f32) if (STORE_Y or plan == "layer") else None
                sums = torch.zeros(B, 4, 2, dtype=torch.float64, device=dev)
                pool = (_p(hx), _p(hn), _p(ax), _p(an)) if last else (None, None, None, None)
                with TIMER.span(f"sa_chain_fwd[{plan}:{l + 1}/{L}>{widths[l]}]" if TIMER.detail else "sa_chain_fwd",
                                B * (12 * N + 4 * N * Cf + 4 * P),
                                2 * B * P * sum(a * c for a, c in zip([Cf + 3] + widths[:l], widths[:l + 1]))
                                if plan == "full" else 2 * B * P * ((Cf + 3) if l == 0 else widths[l - 1]) * widths[l]):
                    if plan == "full":
                        arr = (ctypes.c_int * 3)(*(widths[:l + 1] + [0, 0])[:3])
                        _lib.check(lib.ogc_sa_chain_fwd(
                            B, N, M, S, Cf, 1, l + 1, arr, _p(xyz), _p(new_xyz), _p(feat_pm), _p(idx), None, None,
                            _p(w2d[0]), _p(w2d[1]) if L > 1 else None, _p(w2d[2]) if L > 2 else None,
                            _p(sss[0]) if l > 0 else None, _p(sss[1]) if l > 1 else None, _p(sums), _p(y), *pool, _st()),
                            "ogc_sa_chain_fwd")
                    elif l == 0:
                        arr = (ctypes.c_int * 3)(widths[0], 0, 0)
                        _lib.check(lib.ogc_sa_chain_fwd(
                            B, N, M, S, Cf, 1, 1, arr, _p(xyz), _p(new_xyz), _p(feat_pm), _p(idx), None, None,
                            _p(w2d[0]), None, None, None, None, _p(sums), _p(y), *pool, _st()), "ogc_sa_chain_fwd")
                    else:
                        arr = (ctypes.c_int * 3)(widths[l], 0, 0)
                        _lib.check(lib.ogc_sa_chain_fwd(
                            B, N, M, S, widths[l - 1], 0, 1, arr, None, None, None, None, _p(ys[l - 1]), _p(sss[l - 1]),
                            _p(w2d[l]), None, None, None, None, _p(sums), _p(y) if STORE_Y else None, *pool, _st()),
                            "ogc_sa_chain_fwd")
                gamma, beta = params[3 * l + 1], params[3 * l + 2]
                ss = torch.empty(B, widths[l], 2, **f32)
                mr = torch.empty(B, 4, 2, **f32)
                _lib.check(lib.ogc_gn_finalize(B, widths[l], (widths[l] // 4) * P, _p(sums), _p(gamma.detach()),
                                               _p(beta.detach()), _p(ss), _p(mr), _st()), "ogc_gn_finalize")
                be.launches += 2
                ys.append(y); sss.append(ss); mrs.append(mr)
            out = torch.empty(B, cL, M, **f32)
            sel = torch.empty(B, cL, M, dtype=torch.uint8, device=dev)
            ysel = torch.empty(B, cL, M, **f32)
            _lib.check(lib.ogc_sa_pool_finish(B, cL, M, _p(hx), _p(hn), _p(ax), _p(an), _p(sss[-1]), _p(out), None,
                                              cL, 0, _p(sel), _p(ysel), _st()), "ogc_sa_pool_finish")
            be.launches += 1
            ctx.dims = (B, N, M, S, Cf, L)
            ctx.feat_needs_grad = feat_pm.requires_grad
            ctx.has_feat = True
            ctx.param_objs = params
            ctx.save_for_backward(xyz, new_xyz, feat_pm, idx, sel, ysel, *ys, *sss, *mrs, *[p.detach() for p in params])
            return out
        sums_all = torch.zeros(L, B, 4, 2, dtype=torch.float64, device=dev)      # one fill for all layers
        for l in range(L):
            W, gamma, beta = params[3 * l], params[3 * l + 1], params[3 * l + 2]
            cout, cin = W.shape[0], W.shape[1]
            last = l == L - 1
            y = torch.empty(B, cout, P, **f32)
            sums = sums_all[l]
            if last:
                ymax = torch.empty(B, cout, M, **f32)
                ymin = torch.empty(B, cout, M, **f32)
                amax = torch.empty(B, cout, M, dtype=torch.uint8, device=dev)
                amin = torch.empty(B, cout, M, dtype=torch.uint8, device=dev)
            else:
                ymax = ymin = amax = amin = None
            flops_bytes = B * (4 * cout * P + (4 * cin * P if l else 4 * P + 12 * N + 4 * N * Cf))
            use_tma = l > 0 and _fwd_tma_ok(S, M, cin, cout)
            use_nw = not use_tma and l > 0 and _narrow_ok(S, cin, cout)
            use_tc = not use_tma and not use_nw and _tc_ok(S, cin, cout, l == 0)
            tag = "sa_fwd_tma" if use_tma else "sa_mlp_fwd_nw" if use_nw else "sa_mlp_fwd_tc" if use_tc else "sa_mlp_fwd"
            with TIMER.span(f"{tag}[{cin}>{cout}]" if TIMER.detail else tag, flops_bytes, 2 * B * P * cin * cout):
                if use_tma:
                    w2d = W.detach().reshape(cout, cin).contiguous()
                    _lib.check(lib.ogc_sa_fwd_tma(
                        B, M, S, cin, cout, int(last), _p(y_prev), _p(ss_prev), _p(w2d), _p(y), _p(sums), _p(ymax), _p(ymin),
                        _p(amax), _p(amin), _st()), "ogc_sa_fwd_tma")
                elif use_nw:
                    w2d = W.detach().reshape(cout, cin).contiguous()
                    _lib.check(lib.ogc_sa_mlp_narrow_fwd(
                        B, M, S, cin, cout, int(last), _p(y_prev), _p(ss_prev), _p(w2d), _p(gamma.detach()), _p(y), _p(sums), _p(ymax),
                        _p(ymin), _p(amax), _p(amin), _st()), "ogc_sa_mlp_narrow_fwd")
                elif use_tc:
                    w2d = W.detach().reshape(cout, cin).contiguous()
                    _lib.check(lib.ogc_sa_mlp_layer_fwd_tc(
                        B, N, M, S, cin, cout, int(l == 0), int(last), _p(xyz), _p(new_xyz), _p(feat_pm), _p(idx),
                        _p(y_prev), _p(ss_prev), _p(w2d), _p(y), _p(sums), _p(ymax), _p(ymin), _p(amax), _p(amin),
                        _st()), "ogc_sa_mlp_layer_fwd_tc")
                else:
                    wt = W.detach().reshape(cout, cin).t().contiguous()       # only the SIMT kernel wants W^T
                    _lib.check(lib.ogc_sa_mlp_layer_fwd(
                        B, N, M, S, cin, cout, int(l == 0), int(last), _p(xyz), _p(new_xyz), _p(feat_pm), _p(idx),
                        _p(y_prev), _p(ss_prev), _p(wt), _p(y), _p(sums), _p(ymax), _p(ymin), _p(amax), _p(amin),
                        _st()), "ogc_sa_mlp_layer_fwd")
            ss = torch.empty(B, cout, 2, **f32)
            mr = torch.empty(B, 4, 2, **f32)
            _lib.check(lib.ogc_gn_finalize(B, cout, (cout // 4) * P, _p(sums), _p(gamma.detach()), _p(beta.detach()),
                                           _p(ss), _p(mr), _st()), "ogc_gn_finalize")
            be.launches += 2
            ys.append(y); sss.append(ss); mrs.append(mr)
            y_prev, ss_prev = y, ss
        cout = params[3 * (L - 1)].shape[0]
        out = torch.empty(B, cout, M, **f32)
        sel = torch.empty(B, cout, M, dtype=torch.uint8, device=dev)
        ysel = torch.empty(B, cout, M, **f32)
        _lib.check(lib.ogc_sa_finish(B, cout, M, _p(ymax), _p(ymin), _p(amax), _p(amin), _p(sss[-1]), _p(out), None,
                                     cout, 0, _p(sel), _p(ysel), _st()), "ogc_sa_finish")
        be.launches += 1
        ctx.param_objs = params
        ctx.dims = (B, N, M, S, Cf, L)
        ctx.feat_needs_grad = feat_pm is not None and feat_pm.requires_grad
        ctx.has_feat = feat_pm is not None
        ctx.save_for_backward(xyz, new_xyz, feat_pm if feat_pm is not None else xyz.new_empty(0), idx, sel, ysel,
                              *ys, *sss, *mrs, *[p.detach() for p in params])
        return out

    @staticmethod
    def backward(ctx, go):
        be = get_backend()
        lib = be.lib
        B, N, M, S, Cf, L = ctx.dims
        P = M * S
        saved = ctx.saved_tensors
        xyz, new_xyz, feat_pm, idx, sel, ysel = saved[:6]
        if not ctx.has_feat:
            feat_pm = None
        ys, sss, mrs = saved[6:6 + L], saved[6 + L:6 + 2 * L], saved[6 + 2 * L:6 + 3 * L]
        params = saved[6 + 3 * L:]
        dev = xyz.device
        f32 = dict(dtype=torch.float32, device=dev)
        go = go.contiguous()
        grads = [None] * (3 * L)
        dfeat_pm = None

        cL = params[3 * (L - 1)].shape[0]
        tg = grad_targets(ctx.param_objs)          # accumulate straight into the parameters' .grad (trainer's backward)
        ab_all = torch.zeros(L, B, 4, 2, dtype=torch.float64, device=dev)        # one fill for all layers
        ab = ab_all[L - 1]
        dgamma = tg[3 * (L - 1) + 1] if tg else torch.zeros(cL, **f32)
        dbeta = tg[3 * (L - 1) + 2] if tg else torch.zeros(cL, **f32)
        _lib.check(lib.ogc_sa_last_stats(B, cL, M, _p(go), cL, 0, _p(sel), _p(ysel), _p(mrs[-1]), _p(params[3 * (L - 1) + 1]),
                                         _p(ab), _p(dgamma), _p(dbeta), _st()), "ogc_sa_last_stats")
        be.launches += 1
        dz = None                      # last layer: synthesised from go / sel inside the kernels
        for l in range(L - 1, -1, -1):
            W, gamma = params[3 * l], params[3 * l + 1]
            cout, cin = W.shape[0], W.shape[1]
            w2d = W.reshape(cout, cin).contiguous()
            coef = torch.empty(B, cout, 4, **f32)
            _lib.check(lib.ogc_gn_bwd_coef(B, cout, (cout // 4) * P, _p(ab), _p(mrs[l]), _p(gamma), _p(coef), _st()),
                       "ogc_gn_bwd_coef")
            if not tg:
                grads[3 * l + 1], grads[3 * l + 2] = dgamma, dbeta
            dW = tg[3 * l].view(cout, cin) if tg else torch.zeros(cout, cin, **f32)
            gather = l == 0
            dw_nw = not gather and _narrow_ok(S, cin, cout)
            dw_tc = not dw_nw and _tc_dw_ok(S, cin, cout)
            dw_fn, dw_tag = (lib.ogc_sa_mlp_layer_dw_tc, "sa_mlp_dw_tc") if dw_tc else (lib.ogc_sa_mlp_layer_dw, "sa_mlp_dw")
            if not gather and _dw_tma_ok(S, M, cin, cout):
                with TIMER.span(f"sa_dw_tma[{cin}>{cout}]" if TIMER.detail else "sa_dw_tma", B * P * 4 * (2 * cout + cin), 2 * B * P * cin * cout):
                    _lib.check(lib.ogc_sa_dw_tma(
                        B, M, S, cout, cin, _p(dz), _p(go), cL, 0, _p(sel), _p(ys[l]), _p(coef), _p(ys[l - 1]),
                        _p(sss[l - 1]), _p(dW), _st()), "ogc_sa_dw_tma")
            elif dw_nw:
                with TIMER.span(f"sa_mlp_dw_nw[{cin}>{cout}]" if TIMER.detail else "sa_mlp_dw_nw", B * P * 4 * (2 * cout + cin), 2 * B * P * cin * cout):
                    _lib.check(lib.ogc_sa_mlp_narrow_dw(
                        B, M, S, cout, cin, _p(dz), _p(go), cL, 0, _p(sel), _p(ys[l]), _p(coef), _p(ys[l - 1]),
                        _p(sss[l - 1]), _p(dW), _st()), "ogc_sa_mlp_narrow_dw")
            else:
                with TIMER.span(f"{dw_tag}[{cin}>{cout}]" if TIMER.detail else dw_tag, B * P * 4 * (2 * cout + cin), 2 * B * P * cin * cout):
                    _lib.check(dw_fn(
                        B, N, M, S, cout, cin, int(gather), _p(dz), _p(go), cL, 0, _p(sel), _p(ys[l]), _p(coef),
                        _p(ys[l - 1]) if l else None, _p(sss[l - 1]) if l else None,
                        _p(xyz), _p(new_xyz), _p(feat_pm), _p(idx), _p(dW), _st()), "ogc_sa_mlp_layer_dw")
            be.launches += 2
            if not tg:
                grads[3 * l] = dW.view_as(W)
            if l > 0:
                cprev = params[3 * (l - 1)].shape[0]
                dz_prev = torch.empty(B, cprev, P, **f32)
                ab_prev = ab_all[l - 1]
                dgamma_prev = tg[3 * (l - 1) + 1] if tg else torch.zeros(cprev, **f32)
                dbeta_prev = tg[3 * (l - 1) + 2] if tg else torch.zeros(cprev, **f32)
                dx_nw = _narrow_ok(S, cprev, cout)
                dx_tc = not dx_nw and _tc_dx_ok(S, cout, cprev, False)
                dx_fn, dx_tag = (lib.ogc_sa_mlp_layer_dx_tc, "sa_mlp_dx_tc") if dx_tc else (lib.ogc_sa_mlp_layer_dx, "sa_mlp_dx")
                if dx_nw:
                    with TIMER.span(f"sa_mlp_dx_nw[{cout}>{cprev}]" if TIMER.detail else "sa_mlp_dx_nw", B * P * 4 * (2 * cout + 2 * cprev), 2 * B * P * cout * cprev):
                        _lib.check(lib.ogc_sa_mlp_narrow_dx(
                            B, M, S, cout, cprev, _p(dz), _p(go), cL, 0, _p(sel), _p(ys[l]), _p(coef), _p(w2d),
                            _p(ys[l - 1]), _p(sss[l - 1]), _p(mrs[l - 1]), _p(params[3 * (l - 1) + 1]), _p(dz_prev),
                            _p(ab_prev), _p(dgamma_prev), _p(dbeta_prev), _st()), "ogc_sa_mlp_narrow_dx")
                elif _dx_tma_ok(S, cout, cprev, dz is None):
                    with TIMER.span(f"sa_dx_tma[{cout}>{cprev}]" if TIMER.detail else "sa_dx_tma", B * P * 4 * (2 * cout + 2 * cprev), 2 * B * P * cout * cprev):
                        _lib.check(lib.ogc_sa_dx_tma(
                            B, M, S, cout, cin, 0, cprev, _p(dz), _p(go), cL, 0, _p(sel), _p(ys[l]), _p(coef), _p(w2d),
                            _p(ys[l - 1]), _p(sss[l - 1]), _p(mrs[l - 1]), _p(params[3 * (l - 1) + 1]), _p(dz_prev),
                            _p(ab_prev), _p(dgamma_prev), _p(dbeta_prev), _st()), "ogc_sa_dx_tma")
                    be.launches += 1 if cout > 128 else 0
                elif _chain_dx_ok(S, M, cout, cprev, False) or (cprev % 64 == 0 and _chain_dx_ok(S, M, cout, cprev // 2, False)):
                    # a layer whose resident W^T (hi + lo) exceeds one SM runs as two launches of half the output channels
                    nblk = 1 if _chain_dx_ok(S, M, cout, cprev, False) else 2
                    rows = cprev // nblk
                    chan_sums = torch.zeros(B, cprev, 2, **f32)
                    with TIMER.span(f"sa_chain_dx[{cout}>{cprev}]" if TIMER.detail else "sa_chain_dx", B * P * 4 * (2 * cout + 2 * cprev), 2 * B * P * cout * cprev):
                        for off in range(0, cprev, rows):
                            _lib.check(lib.ogc_sa_chain_dx(
                                B, N, M, S, cout, cin, off, rows, _p(dz), _p(go), cL, 0, _p(sel), _p(ys[l]), _p(coef), _p(w2d),
                                _p(ys[l - 1]), _p(sss[l - 1]), _p(mrs[l - 1]), _p(params[3 * (l - 1) + 1]), _p(dz_prev),
                                _p(ab_prev), _p(dgamma_prev), _p(dbeta_prev), None, None, 0, 0, _p(chan_sums), cprev, off,
                                _st()), "ogc_sa_chain_dx")
                    be.launches += 2 * nblk - 1      # + dx_finalize_kernel per launch
                else:
                    with TIMER.span(f"{dx_tag}[{cout}>{cprev}]" if TIMER.detail else dx_tag, B * P * 4 * (2 * cout + 2 * cprev), 2 * B * P * cout * cprev):
                        _lib.check(dx_fn(
                            B, N, M, S, cout, cin, 0, cprev, _p(dz), _p(go), cL, 0, _p(sel), _p(ys[l]), _p(coef), _p(w2d),
                            _p(ys[l - 1]), _p(sss[l - 1]), _p(mrs[l - 1]), _p(params[3 * (l - 1) + 1]), _p(dz_prev),
                            _p(ab_prev), _p(dgamma_prev), _p(dbeta_prev), None, None, 0, 0, _st()), "ogc_sa_mlp_layer_dx")
                be.launches += 1
                if DEBUG_KEEP is not None:
                    DEBUG_KEEP.append((l, dz_prev, ab_prev, coef))
                dz, ab, dgamma, dbeta = dz_prev, ab_prev, dgamma_prev, dbeta_prev
            elif ctx.feat_needs_grad:
                dfeat_pm = torch.zeros(B, N, Cf, **f32)
                for off in range(0, Cf, 128):
                    rows = min(128, Cf - off)
                    if _chain_dx_ok(S, M, cout, rows, True):
                        with TIMER.span(f"sa_chain_dx[{cout}>scatter{rows}]" if TIMER.detail else "sa_chain_dx", B * P * 4 * (2 * cout + rows), 2 * B * P * cout * rows):
                            _lib.check(lib.ogc_sa_chain_dx(
                                B, N, M, S, cout, cin, 3 + off, rows, _p(dz), _p(go), cL, 0, _p(sel), _p(ys[l]), _p(coef),
                                _p(w2d), None, None, None, None, None, None, None, None, _p(idx), _p(dfeat_pm), Cf, off,
                                None, 0, 0, _st()), "ogc_sa_chain_dx")
                        be.launches += 1
                        continue
                    dx_tc = _tc_dx_ok(S, cout, rows, True)
                    dx_fn, dx_tag = (lib.ogc_sa_mlp_layer_dx_tc, "sa_mlp_dx_tc") if dx_tc else (lib.ogc_sa_mlp_layer_dx, "sa_mlp_dx")
                    with TIMER.span(f"{dx_tag}[{cout}>scatter{rows}]" if TIMER.detail else dx_tag, B * P * 4 * (2 * cout + rows), 2 * B * P * cout * rows):
                        _lib.check(dx_fn(
                            B, N, M, S, cout, cin, 3 + off, rows, _p(dz), _p(go), cL, 0, _p(sel), _p(ys[l]), _p(coef),
                            _p(w2d), None, None, None, None, None, None, None, None, _p(idx), _p(dfeat_pm), Cf, off,
                            _st()), "ogc_sa_mlp_layer_dx")
                    be.launches += 1
        return (None, None, dfeat_pm, None, *grads)


def fused_sa_mlp(xyz, new_xyz, feat_pm, idx, layers):
    """xyz (B,N,3), new_xyz (B,M,3), feat_pm (B,N,Cf) point-major or None, idx (B,M,64) int32,
    layers = [(W, gamma, beta), ...]  ->  (B, C_L, M)."""
    flat = [t for layer in layers for t in layer]
    return _FusedSAMLP.apply(xyz.contiguous(), new_xyz.contiguous(),
                             None if feat_pm is None else feat_pm.contiguous(), idx.contiguous(), *flat)


def supported(nsample, channels):
    """Shapes the fused kernels cover: nsample 64, output widths multiples of 16 up to 256, and the
    (weights + operand tile) of every layer within one CTA's shared memory."""
    if nsample != 64:
        return False
    for cin, cout in zip(channels[:-1], channels[1:]):
        if cout % 16 != 0 or cout > 256:
            return False
        r_t, p_t = (32, 512) if cout <= 32 else (64, 256) if cout <= 64 else (128, 128) if cout <= 128 else (256, 64)
        if cin * (r_t + p_t + 4) * 4 > 225 * 1024:
            return False
    return True
