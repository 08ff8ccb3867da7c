"""torch.profiler view of the training step (eager): CUDA time of the torch (aten) operators by call site."""
import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from ogc_b200 import backend, data
from torch.profiler import profile, ProfilerActivity
dev = torch.device('cuda', 0)
be = backend.get_backend()
tr = bench.build_trainer(dev, 1, 4)
batches = [tuple(x.to(dev) for x in data.make_batch(i, 4, 8192, aug=True, fps_fn=be.fps, device=dev)) for i in range(2)]
for i in range(3): tr.train_step(100000 + i, batches[i % 2], aug_transform=True)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True, record_shapes=True) as prof:
    for i in range(2): tr.train_step(100000 + i, batches[i % 2], aug_transform=True)
    torch.cuda.synchronize()
print(prof.key_averages(group_by_input_shape=True).table(sort_by="cuda_time_total", row_limit=60, max_name_column_width=50, max_shapes_column_width=70))
