"""GPU: per-kernel times of object_aware_icp (64 clouds x 8192 points, 20 iterations)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from ogc_b200 import backend, icp
dev = torch.device("cuda", 0)
be = backend.get_backend()
args = bench.icp_inputs(64, dev)
if args is None:
    import inspect
    fn = [v for k, v in vars(bench).items() if callable(v) and "icp" in k.lower()]
    print([f.__name__ for f in fn]); sys.exit(0)
for _ in range(2): icp.object_aware_icp(*args, icp_iter=20)
torch.cuda.synchronize()
backend.TIMER.enabled = True; backend.TIMER.reset()
icp.object_aware_icp(*args, icp_iter=20)
torch.cuda.synchronize(); backend.TIMER.enabled = False
for k, v in backend.TIMER.summary().items():
    print(f"{k:28s} x{v['calls']:<4d} {v['ms']:8.3f} ms total  {v['ms'] / v['calls']:7.3f} ms each")
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record(); icp.object_aware_icp(*args, icp_iter=20); e.record(); torch.cuda.synchronize()
print("whole call", s.elapsed_time(e), "ms")
