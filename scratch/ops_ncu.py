"""GPU: the operators BASELINE.json's metric names (FPS, ball_query) plus the loss k-NN, at the sizes of the KITTI-SF
step, on KITTI-SF-like scenes -- the target of the `ncu --set full` capture behind profiles/r02_ncu_ops_*.csv.

    ncu --set full --clock-control none -k regex:"fps|ball_query|knn" -o gpurun_out/r02_ops python scratch/ops_ncu.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ogc_b200 import data
from ogc_b200.backend import get_backend

be = get_backend()
dev = torch.device("cuda", 0)
pcs = data.make_batch(7, 4, 8192, aug=True, fps_fn=be.fps, device=dev)[0].to(dev)      # (4,4,8192,3)
flat = pcs.view(16, 8192, 3).contiguous()
view = pcs[:, 0].contiguous()
reps = int(os.environ.get("REPS", "2"))
for _ in range(reps):
    c1 = be.fps(flat, 2048)
    ctr = torch.gather(flat, 1, c1.long()[..., None].expand(-1, -1, 3)).contiguous()
    be.fps(ctr, 1024)
    be.ball_query(2.0, 64, view, view)
    be.knn(64, ctr, flat)
    if hasattr(be, "knn_bounded"):
        be.knn_bounded(32, view, view, 1.0)
    else:
        be.knn(32, view, view)
torch.cuda.synchronize()
print("ok")
