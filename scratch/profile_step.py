import sys, os; sys.path.insert(0,'.')
import torch, json
import bench
from ogc_b200 import backend, data
from torch.profiler import profile, ProfilerActivity
dev=torch.device('cuda',0)
be=backend.get_backend()
tr=bench.build_trainer(dev,1)
batches=[tuple(x.to(dev) for x in data.make_batch(i,4,8192,aug=True,fps_fn=be.fps,device=dev)) for i in range(2)]
for i in range(3): tr.train_step(100000+i,batches[i%2],aug_transform=True)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for i in range(2): tr.train_step(100000+i,batches[i%2],aug_transform=True)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70))
