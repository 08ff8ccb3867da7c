"""Raw-scene FPS (B = 1): cluster kernel vs the single-CTA fallback (OGC_FPS_NO_CLUSTER=1)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ogc_b200 import backend
be = backend.get_backend()
for n in (30000, 60000, 100000, 131072):
    x = (torch.rand(1, n, 3, device="cuda") - 0.5) * 60
    be.fps(x, 64); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); idx = be.fps(x, 8192); e.record(); torch.cuda.synchronize()
    print(f"n={n} m=8192 {'fallback' if os.environ.get('OGC_FPS_NO_CLUSTER') else 'cluster'}: {s.elapsed_time(e):.2f} ms  checksum {int(idx.sum())}")
