"""Multi-GPU bring-up probe: prints a timestamped line after every stage so that a hang can be located.
torchrun --nproc-per-node 2 scratch/nccl_probe.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
t0 = time.time()
rank, lr, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
def say(msg):
    print(f"[{time.time()-t0:7.2f}s r{rank}] {msg}", flush=True)
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
say("init_process_group ...")
dist.init_process_group("nccl", device_id=dev)
say("barrier ...")
dist.barrier()
torch.cuda.synchronize()
x = torch.ones(1024, device=dev) * (rank + 1)
say("eager all_reduce ...")
dist.all_reduce(x)
torch.cuda.synchronize()
say(f"ok sum={float(x[0])}")
import bench
from ogc_b200 import backend, data
be = backend.get_backend()
trainer = bench.build_trainer(dev, world)
batch = tuple(t.to(dev) for t in data.make_batch(1000 * rank, 2, 8192, aug=True, fps_fn=be.fps, device=dev))
for i in range(2):
    say(f"eager step {i} ...")
    d = trainer.train_step(100000 + i, batch, aug_transform=True)
    torch.cuda.synchronize()
say(f"eager ok {d['sum']:.4f}")
for i in range(3):
    say(f"graphed step {i} ...")
    d = trainer.train_step_graphed(100000 + i, batch, aug_transform=True)
    torch.cuda.synchronize()
say(f"graphed ok {d['sum']:.4f}")
p = trainer.opt.flat_p.double().sum()
ps = [torch.zeros_like(p) for _ in range(world)]
dist.all_gather(ps, p)
say(f"param checksums equal across ranks: {all(float(q) == float(ps[0]) for q in ps)}")
dist.destroy_process_group()
say("done")
