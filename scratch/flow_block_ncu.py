"""One FlowStep3D block (flow_regressor.sa1 of configs[2]: 16 samples, 64+3 -> 64 -> 64 -> 64, 512 centres x 16 neighbours)
forward + backward through ogc_b200.bn_fused, for ncu (every launch of the profiled pass is listed):

    ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,... --clock-control none \
        --csv --log-file gpurun_out/flow_block.csv python scratch/flow_block_ncu.py
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn as nn
from ogc_b200 import bn_fused
from ogc_b200.backend import B200Backend, set_backend

set_backend(B200Backend())
torch.manual_seed(0)
B, cin, M, S, widths = 16, 67, 512, 16, [64, 64, 64]
convs, bns = nn.ModuleList(), nn.ModuleList()
last = cin
for c in widths:
    convs.append(nn.Conv2d(last, c, 1, bias=False)); bns.append(nn.BatchNorm2d(c)); last = c
convs, bns = convs.cuda(), bns.cuda()
x = torch.randn(B, cin, M, S, device="cuda", requires_grad=True)
probe = torch.randn(B, widths[-1], M, device="cuda")


def once():
    out = bn_fused.fused_bn_mlp(x, convs, bns)
    (out * probe).sum().backward()


for _ in range(2):
    once()
torch.cuda.synchronize()
torch.cuda.profiler.start()
once()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
