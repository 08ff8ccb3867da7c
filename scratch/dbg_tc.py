import sys; sys.path.insert(0,'.')
import torch
from ogc_b200 import segnet, sa_fused
from ogc_b200.sa_fused import fused_sa_mlp
import pointnet2.pointnet2 as ops
torch.manual_seed(496)
B,S,N,M,Cf,widths=3,64,400,100,96,[64,64,128]
xyz=torch.randn(B,N,3,device='cuda'); new_xyz=xyz[:,:M].contiguous(); feats=torch.randn(B,Cf,N,device='cuda')
mlp=segnet.SharedMLP([Cf+3]+widths).cuda()
with torch.no_grad():
    for n_,p_ in mlp.named_parameters():
        if 'gn.weight' in n_: p_.copy_(torch.randn_like(p_)*0.5+0.8)
        if 'gn.bias' in n_: p_.copy_(torch.randn_like(p_)*0.3)
dist,idx=ops.knn(S,new_xyz,xyz); idx=ops.clip_neighbours_by_radius(dist,idx,1.2)
probe=torch.randn(B,widths[-1],M,device='cuda')
layers=[(getattr(mlp,f"layer{i}").conv.weight,getattr(mlp,f"layer{i}").normlayer.gn.weight,getattr(mlp,f"layer{i}").normlayer.gn.bias) for i in range(3)]
res={}
for tc in (False,True):
    sa_fused.USE_TC=tc
    mlp.zero_grad()
    f=feats.clone().requires_grad_(True)
    out=fused_sa_mlp(xyz,new_xyz,f.transpose(1,2).contiguous(),idx,layers)
    (out*probe).sum().backward()
    res[tc]=(out.detach().clone(), f.grad.clone(), {n:p.grad.clone() for n,p in mlp.named_parameters()})
o0,g0,p0=res[False]; o1,g1,p1=res[True]
rel=lambda a,b: float((a-b).abs().max()/b.abs().max()); mean=lambda a,b: float((a-b).abs().mean()/b.abs().mean())
print("out", rel(o1,o0), mean(o1,o0)); print("dfeat", rel(g1,g0), mean(g1,g0))
for n in p0: print(n, rel(p1[n],p0[n]), mean(p1[n],p0[n]))
# composed reference
mlp.zero_grad(); f=feats.clone().requires_grad_(True)
grouped=torch.cat([ops.grouping_operation(xyz.transpose(1,2).contiguous(),idx)-new_xyz.transpose(1,2).unsqueeze(-1), ops.grouping_operation(f,idx)],1)
ref=mlp(grouped).max(dim=3).values; (ref*probe).sum().backward()
pr={n:p.grad.clone() for n,p in mlp.named_parameters()}
print("vs composed: simt / tc")
for n in p0: print(n, "simt", rel(p0[n],pr[n]), mean(p0[n],pr[n]), " tc", rel(p1[n],pr[n]), mean(p1[n],pr[n]))
print("---- intermediate y tensors: TC vs SIMT")
ys={}
for tc in (False,True):
    sa_fused.USE_TC=tc
    f=feats.clone().requires_grad_(True)
    out=fused_sa_mlp(xyz,new_xyz,f.transpose(1,2).contiguous(),idx,layers)
    sv=out.grad_fn.saved_tensors
    ys[tc]=[t.clone() for t in sv[6:6+9]]
for l in range(3):
    a,b=ys[False][l],ys[True][l]
    ss=ys[False][3+l]
    z0=ss[:,:,0:1]*a+ss[:,:,1:2]; z1=ys[True][3+l][:,:,0:1]*b+ys[True][3+l][:,:,1:2]
    print(f"layer{l}: y max abs diff {float((a-b).abs().max()):.3e} (|y| max {float(a.abs().max()):.2f}); relu sign flips {int(((z0>0)!=(z1>0)).sum())} of {a.numel()}; ss diff {float((ys[False][3+l]-ys[True][3+l]).abs().max()):.2e}")
print("---- all saved tensors")
sv={}
for tc in (False,True):
    sa_fused.USE_TC=tc
    f=feats.clone().requires_grad_(True)
    out=fused_sa_mlp(xyz,new_xyz,f.transpose(1,2).contiguous(),idx,layers)
    sv[tc]=[t.clone() for t in out.grad_fn.saved_tensors]
names=['xyz','new_xyz','feat_pm','idx','sel','ysel','y0','y1','y2','ss0','ss1','ss2','mr0','mr1','mr2']
for i,n in enumerate(names):
    a,b=sv[False][i],sv[True][i]
    if a.dtype==torch.uint8:
        print(n,"mismatch",int((a!=b).sum()),"of",a.numel(), "n255", int((a==255).sum()), int((b==255).sum()))
    else:
        print(n, float((a.float()-b.float()).abs().max()))
print("---- determinism: each mode twice")
def run(tc):
    sa_fused.USE_TC=tc
    mlp.zero_grad()
    f=feats.clone().requires_grad_(True)
    out=fused_sa_mlp(xyz,new_xyz,f.transpose(1,2).contiguous(),idx,layers)
    (out*probe).sum().backward()
    torch.cuda.synchronize()
    return {n:p.grad.clone() for n,p in mlp.named_parameters()}
for tc in (False,True,False,True):
    a=run(tc); b=run(tc)
    print("tc",tc,{n[:6]+n[-6:]: f"{rel(a[n],b[n]):.1e}" for n in a})
a=run(False); b=run(True)
print("simt vs tc again", {n[:6]+n[-6:]: f"{rel(b[n],a[n]):.1e}" for n in a})
