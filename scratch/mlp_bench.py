"""Per-layer timing of the fused SA MLP kernels at KITTI-SF sizes (B=16)."""
import sys; sys.path.insert(0,'.')
import torch
from ogc_b200 import backend, segnet
from ogc_b200.sa_fused import fused_sa_mlp
import pointnet2.pointnet2 as ops
be=backend.get_backend()
B=16
torch.manual_seed(0)
cfgs=[("SA1a",8192,2048,3,[32,32,32]),("SA1b",8192,2048,3,[32,32,64]),("SA2",2048,1024,96,[64,64,128]),("SA3",1024,512,128,[128,128,256])]
reps=int(sys.argv[1]) if len(sys.argv)>1 else 3
import os
if os.environ.get('CFGS'): cfgs=[c for c in cfgs if c[0] in os.environ['CFGS'].split(',')]
from ogc_b200 import sa_fused as _sf
_sf.TC_DW_ALL = os.environ.get("TC_DW_ALL","0")=="1"
for name,N,M,Cf,w in cfgs:
    xyz=(torch.rand(B,N,3,device='cuda')-0.5)*40
    new_xyz=xyz[:,:M].contiguous()
    feat=torch.randn(B,N,Cf,device='cuda',requires_grad=(Cf>3))
    mlp=segnet.SharedMLP([Cf+3]+w).cuda()
    dist,idx=ops.knn(64,new_xyz,xyz)
    layers=[(getattr(mlp,f"layer{i}").conv.weight,getattr(mlp,f"layer{i}").normlayer.gn.weight,getattr(mlp,f"layer{i}").normlayer.gn.bias) for i in range(3)]
    probe=torch.randn(B,w[-1],M,device='cuda')
    for r in range(reps+1):
        if r==1:
            torch.cuda.synchronize(); backend.TIMER.enabled=True; backend.TIMER.detail=True; backend.TIMER.reset()
        out=fused_sa_mlp(xyz,new_xyz,feat,idx,layers)
        (out*probe).sum().backward()
    torch.cuda.synchronize(); backend.TIMER.enabled=False
    P=M*64
    print(f"== {name}: N={N} M={M} P={P} Cin={Cf+3} widths={w}")
    for k,v in backend.TIMER.summary().items():
        if not k.startswith("sa_mlp"): continue
        ms=v['ms']/v['calls']
        import re
        a,bb=re.findall(r'\[(\d+)>(?:scatter)?(\d+)\]',k)[0]; a=int(a); bb=int(bb)
        fl=2*a*bb*P*B
        print(f"   {k:28s} {ms:7.3f} ms   {fl/ms/1e9:7.2f} TFLOP/s   {v['bytes']/v['calls']/ms/1e6:7.1f} GB/s(alg)")
