"""GPU: CUPTI timeline of ONE graphed training step: per-stream busy time, idle gaps of the busiest stream, kernels by time."""
import sys, os, json, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from ogc_b200 import backend, data
from torch.profiler import profile, ProfilerActivity
dev = torch.device('cuda', 0)
be = backend.get_backend()
tr = bench.build_trainer(dev, 1, 4)
batches = [tuple(x.to(dev) for x in data.make_batch(i, 4, 8192, aug=True, fps_fn=be.fps, device=dev)) for i in range(2)]
for i in range(4): tr.train_step_graphed(100000 + i, batches[i % 2], aug_transform=True, next_batch=batches[(i + 1) % 2])
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    tr.train_step_graphed(100010, batches[0], aug_transform=True, next_batch=batches[1])
    torch.cuda.synchronize()
path = "/tmp/step_trace.json"
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "dur" in e]
ev.sort(key=lambda e: e["ts"])
t0 = ev[0]["ts"]; t1 = max(e["ts"] + e["dur"] for e in ev)
print(f"span {(t1 - t0) / 1e3:.3f} ms, {len(ev)} device activities")
by_stream = collections.defaultdict(list)
for e in ev: by_stream[e["args"].get("stream", -1)].append(e)
for s, es in sorted(by_stream.items(), key=lambda kv: -sum(e["dur"] for e in kv[1])):
    print(f"  stream {s}: {len(es)} activities, busy {sum(e['dur'] for e in es) / 1e3:.3f} ms, first at {(es[0]['ts'] - t0) / 1e3:.3f} ms, last end {(max(e['ts'] + e['dur'] for e in es) - t0) / 1e3:.3f} ms")
# union busy time over all streams and global idle gaps
iv = sorted((e["ts"], e["ts"] + e["dur"]) for e in ev)
cur_s, cur_e = iv[0]; busy = 0; gaps = []
for s, e in iv[1:]:
    if s > cur_e:
        busy += cur_e - cur_s; gaps.append((cur_e, s)); cur_s, cur_e = s, e
    else:
        cur_e = max(cur_e, e)
busy += cur_e - cur_s
print(f"device busy (any stream) {busy / 1e3:.3f} ms, idle {(t1 - t0 - busy) / 1e3:.3f} ms in {len(gaps)} gaps; gaps > 15 us:")
for a, b_ in gaps:
    if b_ - a > 15:
        prev = max((e for e in ev if e["ts"] + e["dur"] <= a + 0.5), key=lambda e: e["ts"] + e["dur"])
        nxt = min((e for e in ev if e["ts"] >= b_ - 0.5), key=lambda e: e["ts"])
        print(f"    {(a - t0) / 1e3:8.3f} ms  {b_ - a:6.0f} us   after {prev['name'][:50]}  before {nxt['name'][:50]}")
agg = collections.defaultdict(lambda: [0, 0.0])
for e in ev:
    agg[e["name"][:70]][0] += 1; agg[e["name"][:70]][1] += e["dur"]
print("top activities by time:")
for n, (c, d) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print(f"    {d / 1e3:7.3f} ms  x{c:<4d} {n}")
own = sum(d for n, (c, d) in agg.items() if ("ogc" in n or "chain" in n or "_kernel" in n and "at::" not in n))
print(f"concurrency: sum of activity durations {sum(e['dur'] for e in ev) / 1e3:.3f} ms")
# time where only small (<= 32 CTAs) kernels run is not visible here; list the long single-stream stretches at the start
print("first 40 activities:")
for e in ev[:40]:
    print(f"    {(e['ts'] - t0) / 1e3:7.3f} +{e['dur'] / 1e3:6.3f} ms  s{e['args'].get('stream')}  {e['name'][:70]}")
