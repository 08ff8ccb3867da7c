"""GPU: a 3-launch chain (synth dense -> dense -> scatter) of ogc_sa_chain_dx without host syncs in between, vs the per-layer kernels."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ogc_b200 import backend
be = backend.get_backend(); lib = be.lib
_p = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
st = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
sync_between = os.environ.get("SYNC", "0") == "1"
torch.manual_seed(0)
f32 = dict(dtype=torch.float32, device="cuda")
B, N, M, C = 3, 1024, 512, 128
P = M * 64
T = dict(y2=torch.randn(B, C, P, **f32), y1=torch.randn(B, C, P, **f32), y0=torch.randn(B, C, P, **f32), go=torch.randn(B, C, M, **f32),
         sel=torch.randint(0, 64, (B, C, M), dtype=torch.uint8, device="cuda"),
         coef=[torch.randn(B, C, 4, **f32) * 0.5 for _ in range(3)], W2=torch.randn(C, C, **f32) * 0.1, W1=torch.randn(C, C, **f32) * 0.1,
         W0=torch.randn(C, C + 3, **f32) * 0.1, ss=[torch.randn(B, C, 2, **f32) for _ in range(2)],
         mr=[torch.rand(B, 4, 2, **f32) + 0.5 for _ in range(2)], gamma=[torch.randn(C, **f32) for _ in range(2)],
         idx=torch.randint(0, N, (B, M, 64), dtype=torch.int32, device="cuda"))


def chain(new):
    fn = lib.ogc_sa_chain_dx if new else lib.ogc_sa_mlp_layer_dx_tc
    outs = []
    dz = None
    ys = [T["y0"], T["y1"], T["y2"]]
    Ws = [T["W0"], T["W1"], T["W2"]]
    for l in (2, 1):
        dz_prev = torch.empty(B, C, P, **f32); ab = torch.zeros(B, 4, 2, dtype=torch.float64, device="cuda")
        dg = torch.zeros(C, **f32); db = torch.zeros(C, **f32); cs = torch.zeros(B, C, 2, **f32)
        args = [B, N, M, 64, C, C, 0, C, _p(dz), _p(T["go"]), C, 0, _p(T["sel"]), _p(ys[l]), _p(T["coef"][l]), _p(Ws[l]),
                _p(ys[l - 1]), _p(T["ss"][l - 1]), _p(T["mr"][l - 1]), _p(T["gamma"][l - 1]), _p(dz_prev), _p(ab), _p(dg), _p(db), None, None, 0, 0]
        if new:
            args += [_p(cs), 0, 0]
        assert fn(*args, st()) == 0
        if sync_between:
            torch.cuda.synchronize()
        outs += [dz_prev, ab.float(), dg, db]
        dz = dz_prev
    dfeat = torch.zeros(B, N, C, **f32)
    args = [B, N, M, 64, C, C + 3, 3, C, _p(dz), _p(T["go"]), C, 0, _p(T["sel"]), _p(ys[0]), _p(T["coef"][0]), _p(Ws[0])] + [None] * 8 + \
           [_p(T["idx"]), _p(dfeat), C, 0]
    if new:
        args += [None, 0, 0]
    assert fn(*args, st()) == 0
    torch.cuda.synchronize()
    return outs + [dfeat]


ref = chain(False)
names = ["dz1", "ab1", "dg1", "db1", "dz0", "ab0", "dg0", "db0", "dfeat"]
bad = 0
for r in range(reps):
    out = chain(True)
    rels = [float((a - b_).norm() / a.norm()) for a, b_ in zip(ref, out)]
    if max(rels) > 1e-5:
        bad += 1
        if bad <= 4:
            print("  bad run", r, {n: f"{x:.1e}" for n, x in zip(names, rels) if x > 1e-5})
            k = [i for i, x in enumerate(rels) if x > 1e-5][0]
            if ref[k].dim() == 3:
                d = (ref[k] - out[k]).abs()
                nz = (d > 1e-3 * ref[k].abs().max()).nonzero()
                print("     first tensor", names[k], "n bad", len(nz), "first", nz[:3].tolist(), "last", nz[-2:].tolist(),
                      "samples", sorted(set(nz[:, 0].tolist())), "chan range", int(nz[:, 1].min()), int(nz[:, 1].max()),
                      "pos range", int(nz[:, 2].min()), int(nz[:, 2].max()))
print(f"sync_between={sync_between}: {bad}/{reps} bad chains")
