"""torch.profiler view of the FlowStep3D training step (eager)."""
import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from torch.profiler import profile, ProfilerActivity
os.environ["OGC_FLOW_EAGER"] = "1"
step, _ = bench.flow_step_fn(2048, 16, 4, torch.device("cuda", 0))
for i in range(3): step(i)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for i in range(2): step(i)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="self_cuda_time_total", row_limit=40, max_name_column_width=80))
