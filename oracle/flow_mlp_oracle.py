"""TEST INFRASTRUCTURE -- CPU restatement (numpy, fp64) of the shared MLP of the FlowStep3D blocks, forward and backward.

Only tests/ may import this module.  It restates, loop for loop, what the reference computes in
utils/flowstep3d_util.py:126-137 (PointNetSetAbstraction.forward) and :52-64 (FlowEmbedding.forward):

    for conv, bn in zip(mlp_convs, mlp_bns):  x = relu(bn(conv(x)))        # Conv2d 1x1 (bias=False), BatchNorm2d in
    out = torch.max(x, -1)[0]                                             # training mode, max over nsample

and the gradient of that expression, in the form csrc/bn_mlp.cu + the contraction kernels evaluate it:

    y_l = W_l a_{l-1}                          a_0 = x,  a_l = relu(s_l y_l + t_l)
    mean_l, var_l over (batch, centre, slot) per channel (biased),  r_l = 1/sqrt(var_l + eps)
    s_l = gamma_l r_l,  t_l = beta_l - mean_l s_l
    running_mean <- (1-m) running_mean + m mean,  running_var <- (1-m) running_var + m var n/(n-1)
    backward, given dz_l = d loss / d (s_l y_l + t_l) already masked by the ReLU:
        A = sum dz_l,  Bx = sum dz_l yhat_l        (yhat = (y - mean) r),      dgamma_l = Bx,  dbeta_l = A
        dY_l = gamma_l r_l (dz_l - A/n - yhat_l Bx/n)
        dW_l = dY_l a_{l-1}^T,   dz_{l-1} = relu'(.) W_l^T dY_l

Parity status: pinned against torch's own Conv2d / BatchNorm2d / ReLU / max autograd on CPU in fp64
(tests/test_oracle.py::test_flow_mlp_oracle_matches_torch_autograd) -- the reference's blocks ARE those torch modules,
so this is the reference's implementation, not a third restatement.
"""
import numpy as np

EPS = 1e-5          # nn.BatchNorm2d default (utils/flowstep3d_util.py:30, :95)


def forward(x, weights, gammas=None, betas=None, running=None, momentum=0.1):
    """x (B,Cin,M,S); weights [W_l (Cout,Cin)]; gammas / betas [(Cout,)] or None for a bare convolution block
    (use_act=False); running = [(running_mean, running_var)] updated in place.  Returns (out (B,C_L,M), cache)."""
    x = np.asarray(x, dtype=np.float64)
    act = gammas is not None
    a, cache = x, {"x": x, "layers": []}
    for l, W in enumerate(weights):
        W = np.asarray(W, dtype=np.float64)
        y = np.einsum("oc,bcms->boms", W, a)
        if act:
            n = y.shape[0] * y.shape[2] * y.shape[3]
            mean = y.mean(axis=(0, 2, 3))
            var = y.var(axis=(0, 2, 3))                      # biased
            r = 1.0 / np.sqrt(var + EPS)
            s = np.asarray(gammas[l], dtype=np.float64) * r
            t = np.asarray(betas[l], dtype=np.float64) - mean * s
            if running is not None:
                rm, rv = running[l]
                rm[...] = (1 - momentum) * rm + momentum * mean
                rv[...] = (1 - momentum) * rv + momentum * var * n / max(n - 1, 1)
            z = s[None, :, None, None] * y + t[None, :, None, None]
            cache["layers"].append({"W": W, "a_prev": a, "y": y, "mean": mean, "r": r, "z": z, "gamma": np.asarray(gammas[l], np.float64)})
            a = np.maximum(z, 0.0)
        else:
            cache["layers"].append({"W": W, "a_prev": a, "y": y})
            a = y
    sel = a.argmax(axis=-1)                                  # first maximum, as the kernel's scan
    cache["sel"], cache["act"] = sel, act
    return np.take_along_axis(a, sel[..., None], axis=-1)[..., 0], cache


def backward(go, cache):
    """go (B,C_L,M) -> (dx (B,Cin,M,S), [dW_l], [dgamma_l], [dbeta_l])."""
    go = np.asarray(go, dtype=np.float64)
    layers, act = cache["layers"], cache["act"]
    S = cache["x"].shape[-1]
    da = np.zeros(layers[-1]["y"].shape)
    np.put_along_axis(da, cache["sel"][..., None], go[..., None], axis=-1)       # pooled gradient at the winning slot
    dWs, dgs, dbs = [], [], []
    for L in reversed(layers):
        if act:
            dz = da * (L["z"] > 0)
            n = dz.shape[0] * dz.shape[2] * dz.shape[3]
            yhat = (L["y"] - L["mean"][None, :, None, None]) * L["r"][None, :, None, None]
            A, Bx = dz.sum(axis=(0, 2, 3)), (dz * yhat).sum(axis=(0, 2, 3))
            dgs.append(Bx); dbs.append(A)
            k = (L["gamma"] * L["r"])[None, :, None, None]
            dY = k * (dz - A[None, :, None, None] / n - yhat * Bx[None, :, None, None] / n)
        else:
            dY = da
        dWs.append(np.einsum("boms,bcms->oc", dY, L["a_prev"]))
        da = np.einsum("oc,boms->bcms", L["W"], dY)
    assert S == da.shape[-1]
    return da, dWs[::-1], dgs[::-1], dbs[::-1]
