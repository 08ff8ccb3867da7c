"""GPU: per-role cycle sums of one CTA of ogc_sa_chain_dx (producer / MMA issuer / epilogue) for the SA2 / SA3 shapes."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ogc_b200 import backend, segnet, sa_fused
import pointnet2.pointnet2 as ops
be = backend.get_backend(); lib = be.lib
B = 16
cfgs = {"SA2": (2048, 1024, 96, [64, 64, 128]), "SA3": (1024, 512, 128, [128, 128, 128])}
for name in (sys.argv[1:] or ["SA2", "SA3"]):
    N, M, Cf, w = cfgs[name]
    torch.manual_seed(0)
    xyz = (torch.rand(B, N, 3, device="cuda") - 0.5) * 40
    new_xyz = xyz[:, :M].contiguous()
    feat = torch.randn(B, N, Cf, device="cuda", requires_grad=True)
    mlp = segnet.SharedMLP([Cf + 3] + w).cuda()
    dist, idx = ops.knn(64, new_xyz, xyz)
    layers = [(getattr(mlp, f"layer{i}").conv.weight, getattr(mlp, f"layer{i}").normlayer.gn.weight, getattr(mlp, f"layer{i}").normlayer.gn.bias) for i in range(3)]
    probe = torch.randn(B, w[-1], M, device="cuda")
    sa_fused.USE_CHAIN_DX = True
    for r in range(2):
        out = sa_fused.fused_sa_mlp(xyz, new_xyz, feat, idx, layers)
        dbg = torch.zeros(8 * 32, dtype=torch.int64, device="cuda")
        lib.ogc_sa_chain_dx_debug(ctypes.c_void_p(dbg.data_ptr()))
        (out * probe).sum().backward()
        torch.cuda.synchronize()
        lib.ogc_sa_chain_dx_debug(None)
    d = dbg.view(8, 32).cpu().tolist()
    print("==", name, w)
    for k in range(3):
        r = d[k]
        if r[11] == 0 and r[3] == 0:
            continue
        print(f"  launch {k}: tiles/CTA {r[11]}")
        print(f"    producer (1 of 8 warps): stage-full wait {r[0]}  TMEM slot wait {r[1]}  rebuild+store {r[2]}  total {r[3]}")
        print(f"    issuer: accumulator wait {r[4]}  operand wait {r[5]}  issue {r[6]}  total {r[7]}")
        print(f"    epilogue (1 of 8 warps): accumulator wait {r[8]}  stage-full wait {r[12]}  work {r[9]} = tmem ld {r[13]} + mask/store {r[14]} + transposes/sums {r[15]} + fence/arrive {r[16]}  total {r[10]}")
