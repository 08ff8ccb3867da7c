import sys; sys.path.insert(0,'.')
import torch
from ogc_b200.backend import B200Backend
b=B200Backend()
torch.manual_seed(0)
q=torch.randn(2,100,3).cuda(); r=torch.randn(2,257,3).cuda()
d2,_=b.knn(1,q,r); ds,_=b.knn(1,q,r,sqrt=True)
g=torch.sqrt(d2); c=torch.sqrt(d2.cpu())
print("fused vs torch.cuda.sqrt equal:", torch.equal(ds,g), (ds!=g).sum().item())
print("torch cuda vs cpu sqrt equal:", torch.equal(g.cpu(),c), (g.cpu()!=c).sum().item())
print("fused vs cpu:", (ds.cpu()!=c).sum().item())
bad=(ds.cpu()!=c).nonzero()[:3]
for i in bad: 
    i=tuple(i.tolist()); print(d2[i].item().hex() if hasattr(d2[i].item(),'hex') else d2[i].item(), ds[i].item(), c[i].item())
import numpy as np
x=d2.cpu().numpy(); print("numpy sqrt vs torch cpu:", (np.sqrt(x)!=c.numpy()).sum(), "numpy vs fused", (np.sqrt(x)!=ds.cpu().numpy()).sum())
