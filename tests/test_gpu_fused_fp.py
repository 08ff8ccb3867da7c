"""GPU: the fused feature-propagation block (csrc/fp_mlp.cu + dense kernels of mlp_bwd.cu) against the
torch-composed path of the reference FP module (three_nn -> weights -> three_interpolate -> cat ->
conv1x1 as fp32 matmul -> GroupNorm -> ReLU) on the same device: output and every gradient."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel_err(a, b):
    a, b = a.detach(), b.detach()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))


def fro_err(a, b):
    return float((a - b).norm() / b.norm().clamp_min(1e-12))


@pytest.mark.parametrize("n,m,c1,c2,widths", [
    (1024, 256, 3, 64, [64, 64, 64]),        # kitti FP[0]: skip = raw coordinates (no gradient), 3 layers
    (500, 130, 96, 128, [64, 64]),           # kitti FP[1], ragged n (not a multiple of 4)
    (256, 128, 128, 256, [128, 128]),        # kitti FP[2]: 384 input channels, three row blocks of dX
    (256, 64, 0, 64, [128, 64]),             # no skip features
    (300, 100, 192, 256, [128, 256]),        # 448 inputs, 256-wide output layer
    (200, 64, 192, 256, [256, 128]),         # sapien / ogcdr FP[1]: 256-wide HIDDEN layer (dense dX in two 128-row blocks)
    (512, 256, 3, 128, [128, 128, 64]),      # sapien / ogcdr FP[0]
])
def test_fused_fp_matches_composed(b200, n, m, c1, c2, widths):
    from ogc_b200 import segnet
    torch.manual_seed(n + c1)
    B, dev = 3, "cuda"
    unknown = torch.randn(B, n, 3, device=dev)
    known = unknown[:, torch.randperm(n, device=dev)[:m]].contiguous() + 0.01 * torch.randn(B, m, 3, device=dev)
    known[:, :8] = unknown[:, :8]                           # exact coincidences: dist = 0 -> weight ~ 1
    skip = torch.randn(B, c1, n, device=dev) if c1 else None
    kf = torch.randn(B, c2, m, device=dev).relu()
    fp = segnet.FeaturePropagation([c2 + c1] + widths).to(dev)
    with torch.no_grad():
        for name, p in fp.named_parameters():
            if "gn.weight" in name:
                p.copy_(torch.randn_like(p) * 0.5 + 0.8)
            if "gn.bias" in name:
                p.copy_(torch.randn_like(p) * 0.3)
    probe = torch.randn(B, widths[-1], n, device=dev)

    def run(composed):
        segnet.FORCE_COMPOSED = composed
        try:
            s = skip.clone().requires_grad_(c1 > 3) if skip is not None else None
            k = kf.clone().requires_grad_(True)
            fp.zero_grad()
            out = fp(unknown, known, s, k)
            (out * probe).sum().backward()
            return out, k.grad, (s.grad if s is not None and s.requires_grad else None), \
                {nm: p.grad.clone() for nm, p in fp.named_parameters()}
        finally:
            segnet.FORCE_COMPOSED = False

    ref, ref_dk, ref_ds, ref_g = run(True)
    out, dk, ds, g = run(False)
    assert out.shape == ref.shape
    assert rel_err(out, ref) < 2e-5, rel_err(out, ref)
    # ReLU decisions at z ~ 0 can flip under 1e-6 forward differences (see test_gpu_fused_sa): Frobenius tight, max loose
    assert fro_err(dk, ref_dk) < 2e-3 and rel_err(dk, ref_dk) < 3e-2, (fro_err(dk, ref_dk), rel_err(dk, ref_dk))
    if ref_ds is not None:
        assert fro_err(ds, ref_ds) < 2e-3 and rel_err(ds, ref_ds) < 3e-2, (fro_err(ds, ref_ds), rel_err(ds, ref_ds))
    for nm in ref_g:
        assert fro_err(g[nm], ref_g[nm]) < 2e-3 and rel_err(g[nm], ref_g[nm]) < 3e-2, (nm, fro_err(g[nm], ref_g[nm]))


def test_fused_fp_is_the_path_the_model_takes(b200):
    """MaskFormer3D on the GPU routes its FP levels through libogc_b200 (no cuBLAS conv / torch GroupNorm)."""
    from ogc_b200 import backend, segnet
    torch.manual_seed(0)
    net = segnet.MaskFormer3D(n_slot=6, n_point=512, variant="kitti").cuda()
    pc = torch.randn(2, 512, 3, device="cuda")
    backend.TIMER.enabled = True
    backend.TIMER.reset()
    try:
        net(pc, pc).sum().backward()
        torch.cuda.synchronize()
        names = set(backend.TIMER.summary())
    finally:
        backend.TIMER.enabled = False
        backend.TIMER.reset()
    assert {"fp_interp_concat", "fp_mlp_fwd", "fp_mlp_dw", "fp_mlp_dx"} <= names, names


@pytest.mark.parametrize("n,k", [(1000, 10), (512, 8), (77, 16)])
def test_mask_head_matches_composed(b200, n, k):
    """csrc/mask_head.cu vs normalize -> einsum -> /0.05 -> softmax (models/segnet_kitti.py:85-88), values and both gradients."""
    import torch.nn.functional as F
    from ogc_b200.segnet import _MaskHeadFn
    torch.manual_seed(n)
    B, D = 3, 64
    feats = torch.randn(B, D, n, device="cuda") * torch.rand(B, 1, n, device="cuda") * 3
    feats[:, :, 5] = 0.0                                   # a zero feature vector: F.normalize's eps branch
    slot = torch.randn(B, D, k, device="cuda")
    probe = torch.randn(B, n, k, device="cuda")

    f1, s1 = feats.clone().requires_grad_(True), slot.clone().requires_grad_(True)
    ref = (torch.einsum("bdn,bdk->bnk", F.normalize(f1, dim=1), F.normalize(s1, dim=1)) / 0.05).softmax(dim=-1)
    (ref * probe).sum().backward()
    f2, s2 = feats.clone().requires_grad_(True), slot.clone().requires_grad_(True)
    out = _MaskHeadFn.apply(f2, F.normalize(s2, dim=1), 1.0 / 0.05)
    (out * probe).sum().backward()
    assert float((out.detach() - ref.detach()).abs().max()) < 2e-5
    live = torch.ones(n, dtype=torch.bool, device="cuda"); live[5] = False      # d/df at f = 0 is eps-scaled noise in both
    assert fro_err(f2.grad[:, :, live], f1.grad[:, :, live]) < 1e-4, fro_err(f2.grad[:, :, live], f1.grad[:, :, live])
    assert fro_err(s2.grad, s1.grad) < 1e-4, fro_err(s2.grad, s1.grad)
