"""GPU: which Python lines emit the small torch kernels of a training step (eager, torch.profiler with stacks)."""
import sys, os, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
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
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True) as prof:
    tr.train_step(100005, batches[0], aug_transform=True)
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for e in prof.key_averages(group_by_stack_n=12):
    t = getattr(e, "self_device_time_total", 0) or getattr(e, "self_cuda_time_total", 0)
    if t <= 0 or not e.key.startswith("aten::"):
        continue
    frame = next((f for f in (e.stack or []) if "ogc_b200/" in f or "pointnet2/" in f), "autograd / other")
    frame = frame.replace(root + "/", "")
    agg[(frame[:100], e.key)][0] += e.count
    agg[(frame[:100], e.key)][1] += t
tot = sum(v[1] for v in agg.values())
print(f"aten ops with own device time: {sum(v[0] for v in agg.values())} calls, {tot / 1e3:.3f} ms")
byline = collections.defaultdict(lambda: [0, 0.0])
for (frame, name), (c, t) in agg.items():
    byline[frame][0] += c; byline[frame][1] += t
byop = collections.defaultdict(lambda: [0, 0.0])
for (frame, name), (c, t) in agg.items():
    byop[name][0] += c; byop[name][1] += t
for name, (c, t) in sorted(byop.items(), key=lambda kv: -kv[1][1])[:30]:
    print(f"   {t / 1e3:7.3f} ms x{c:<4d} {name}")
ex = next((e for e in prof.key_averages(group_by_stack_n=12) if e.key == "aten::copy_"), None)
print("example stack:", ex.stack[:12] if ex is not None else None)
for frame, (c, t) in sorted(byline.items(), key=lambda kv: -kv[1][1])[:45]:
    ops = sorted(((n, v) for (f, n), v in agg.items() if f == frame), key=lambda kv: -kv[1][1])[:3]
    print(f"{t / 1e3:7.3f} ms x{c:<4d} {frame}   [{', '.join(f'{n[6:]} x{v[0]}' for n, v in ops)}]")
