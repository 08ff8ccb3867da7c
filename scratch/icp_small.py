"""GPU: object_aware_icp at the reference's own chunk size (4 clouds) and at 8 / 16 clouds."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from ogc_b200 import icp
dev = torch.device("cuda", 0)
for B in (1, 4, 8, 16, 64):
    args = bench.icp_inputs(B, dev)
    for _ in range(2): icp.object_aware_icp(*args, icp_iter=20)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(3): icp.object_aware_icp(*args, icp_iter=20)
    e.record(); torch.cuda.synchronize()
    print(f"B={B}: {s.elapsed_time(e) / 3:.2f} ms per call, {B / (s.elapsed_time(e) / 3e3):.0f} clouds/s")
