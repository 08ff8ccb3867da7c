"""GPU: pin the tcgen05 / TMEM conventions of csrc/tcgen05.cuh on hardware (descriptor bit layout,
128B-swizzled K-major and MN-major operand tiles, TMEM read-back, 3xTF32 accuracy) with a single-CTA GEMM."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu


def run(mode, n, k, split3, a, b):
    import os
    lib = ctypes.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "libogc_probe.so"))
    d = torch.full((128, n), float("nan"), device="cuda")
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    rc = lib.ogc_tc_probe_gemm(mode, n, k, split3, P(a), P(b), P(d), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0, rc
    torch.cuda.synchronize()
    return d


@pytest.mark.parametrize("n,k", [(32, 32), (64, 32), (128, 64), (256, 128), (96, 96)])
def test_k_major_exact_on_tf32_representable_inputs(n, k):
    torch.manual_seed(n + k)
    a = torch.randint(-8, 9, (128, k), device="cuda").float()
    b = torch.randint(-8, 9, (n, k), device="cuda").float()
    d = run(0, n, k, 0, a, b)
    assert torch.equal(d, a @ b.t())


@pytest.mark.parametrize("n,k", [(32, 32), (64, 64), (128, 32), (256, 64), (160, 96)])
def test_mn_major_exact_on_tf32_representable_inputs(n, k):
    torch.manual_seed(n * 3 + k)
    a = torch.randint(-8, 9, (k, 128), device="cuda").float()
    b = torch.randint(-8, 9, (k, n), device="cuda").float()
    d = run(1, n, k, 0, a, b)
    assert torch.equal(d, a.t() @ b)


@pytest.mark.parametrize("mode", [0, 1])
def test_3xtf32_is_fp32_grade(mode):
    torch.manual_seed(5)
    n, k = 128, 96          # (128 + n) * k * 4 B * 2 (hi + lo tiles) must fit one CTA's shared memory
    a = torch.randn((128, k) if mode == 0 else (k, 128), device="cuda")
    b = torch.randn((n, k) if mode == 0 else (k, n), device="cuda")
    ref = (a.double() @ b.double().t()) if mode == 0 else (a.double().t() @ b.double())
    one = run(mode, n, k, 0, a, b).double()
    three = run(mode, n, k, 1, a, b).double()
    fp32 = ((a @ b.t()) if mode == 0 else (a.t() @ b)).double()
    e1 = float((one - ref).abs().max())
    e3 = float((three - ref).abs().max())
    e32 = float((fp32 - ref).abs().max())
    assert e1 > 1e-3          # single-pass TF32 is NOT accurate enough for the 1e-4 logit parity ...
    print(f'single-pass {e1:.2e}  3xTF32 {e3:.2e}  fp32 {e32:.2e}')
    assert e3 < 4 * e32 + 1e-6   # ... the 3-pass split is at the level of an fp32 GEMM (measured: 2.6-3.1x cuBLAS sgemm's error)


@pytest.mark.parametrize("n,k", [(32, 32), (64, 128), (64, 96), (128, 64)])
def test_a_operand_from_tensor_memory(n, k):
    """The .ts form: A written to TMEM with tcgen05.st (lane = row, column = k), B K-major in shared memory."""
    torch.manual_seed(n + 7 * k)
    a = torch.randint(-8, 9, (128, k), device="cuda").float()
    b = torch.randint(-8, 9, (n, k), device="cuda").float()
    assert torch.equal(run(2, n, k, 0, a, b), a @ b.t())
    a, b = torch.randn(128, k, device="cuda"), torch.randn(n, k, device="cuda")
    ref = a.double() @ b.double().t()
    e3 = float((run(2, n, k, 1, a, b).double() - ref).abs().max())
    e32 = float(((a @ b.t()).double() - ref).abs().max())
    assert e3 < 4 * e32 + 1e-6, (e3, e32)


def test_tf32_operands_are_truncated():
    """tcgen05.mma kind::tf32 ignores the 13 low mantissa bits of an fp32 operand (truncation toward zero) on every
    operand path (shared-memory A, shared-memory B, tensor-memory A).  csrc/sa_dw_tma.cu relies on it: the raw fp32
    value serves as the 'hi' operand and x - trunc(x) as the exact 'lo' operand of the 3xTF32 split."""
    n, k = 32, 32
    u = 2.0 ** -10
    fr = torch.tensor([0.0, 0.25, 0.49, 0.5, 0.51, 0.75, 0.999], device="cuda")
    for sign in (1.0, -1.0):
        vals = sign * (1.0 + fr * u)
        a = torch.zeros(128, k, device="cuda"); a[:len(fr), 0] = vals
        b = torch.zeros(n, k, device="cuda"); b[0, 0] = 1.0
        assert torch.equal(run(0, n, k, 0, a, b)[:len(fr), 0], torch.full_like(fr, sign))
        assert torch.equal(run(2, n, k, 0, a, b)[:len(fr), 0], torch.full_like(fr, sign))
        a2 = torch.zeros(128, k, device="cuda"); a2[0, 0] = 1.0
        b2 = torch.zeros(n, k, device="cuda"); b2[:len(fr), 0] = vals
        assert torch.equal(run(0, n, k, 0, a2, b2)[0, :len(fr)], torch.full_like(fr, sign))
