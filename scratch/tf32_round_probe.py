"""GPU: what does tcgen05.mma kind::tf32 do with the 13 low mantissa bits of an fp32 operand: truncate or round?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from test_gpu_tcgen05 import run
n, k = 32, 32
u = 2.0 ** -10
fr = torch.tensor([0.0, 0.25, 0.49, 0.5, 0.51, 0.75, 0.999], device="cuda")
for sign in (1.0, -1.0):
    a = torch.zeros(128, k, device="cuda")
    a[:len(fr), 0] = sign * (1.0 + fr * u)
    b = torch.zeros(n, k, device="cuda"); b[0, 0] = 1.0
    d = run(0, n, k, 0, a, b)[:len(fr), 0]
    print("A operand (smem):", [(f"{float(f):.3f}", float((x * sign - 1.0) / u)) for f, x in zip(fr, d)])
    a2 = torch.zeros(128, k, device="cuda"); a2[0, 0] = 1.0
    b2 = torch.zeros(n, k, device="cuda"); b2[:len(fr), 0] = sign * (1.0 + fr * u)
    d2 = run(0, n, k, 0, a2, b2)[0, :len(fr)]
    print("B operand (smem):", [(f"{float(f):.3f}", float((x * sign - 1.0) / u)) for f, x in zip(fr, d2)])
    d3 = run(2, n, k, 0, a, b)[:len(fr), 0]
    print("A operand (tmem):", [(f"{float(f):.3f}", float((x * sign - 1.0) / u)) for f, x in zip(fr, d3)])
