"""Quick per-op timing at KITTI-SF sizes: ours vs the reference CUDA extension (if built)."""
import sys, time; sys.path.insert(0,'.')
import torch, numpy as np
from ogc_b200.backend import B200Backend
from oracle import refext
b=B200Backend(); r=refext.RefExtBackend() if refext.available() else None
def scene(B,N,seed=0):
    rng=np.random.default_rng(seed)
    return torch.from_numpy((rng.random(size=(B,N,3))*np.array([50,4,30])-np.array([25,2,-5])).astype(np.float32)).cuda()
def t(fn, it=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); s=torch.cuda.Event(enable_timing=True); e=torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(it): fn()
    e.record(); torch.cuda.synchronize(); return s.elapsed_time(e)/it
pc=scene(16,8192); 
sub=pc[:, :2048].contiguous()
pc4=scene(4,8192,1)
rows=[("fps 16x8192->2048", lambda be: be.fps(pc,2048)),
      ("fps 16x2048->1024", lambda be: be.fps(sub,1024)),
      ("knn64 16x(2048 q, 8192 ref)", lambda be: be.knn(64,sub,pc)),
      ("knn64 16x(1024 q, 2048 ref)", lambda be: be.knn(64,sub[:, :1024].contiguous(),sub)),
      ("knn32 4x(8192,8192)", lambda be: be.knn(32,pc4,pc4)),
      ("knn1 4x(8192,8192)", lambda be: be.knn(1,pc4,pc4)),
      ("three_nn 16x(8192 q,2048 ref)", lambda be: be.three_nn(pc,sub)),
      ("ballq r2 s64 4x(8192,8192)", lambda be: be.ball_query(2.0,64,pc4,pc4)),
     ]
for name,fn in rows:
    a=t(lambda: fn(b)); c=t(lambda: fn(r),3) if r else float('nan')
    print(f"{name:34s} ours {a:8.3f} ms   ref {c:9.3f} ms   x{c/a:6.1f}")
f=torch.randn(16,128,2048,device='cuda'); idx=torch.randint(0,2048,(16,1024,64),device='cuda',dtype=torch.int32)
go=torch.randn(16,128,1024,64,device='cuda')
for name,fn in [("group 16x128x(1024x64)", lambda be: be.group_points(f,idx)), ("group_grad", lambda be: be.group_points_grad(go,idx,2048))]:
    a=t(lambda: fn(b)); c=t(lambda: fn(r),3) if r else float('nan')
    gb=(16*128*1024*64*4)/1e9
    print(f"{name:34s} ours {a:8.3f} ms ({gb/a*1e3:7.1f} GB/s)  ref {c:9.3f} ms   x{c/a:6.1f}")
