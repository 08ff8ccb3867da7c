"""GPU: ogc_sa_chain_dx against ogc_sa_mlp_layer_dx_tc on synthetic single-layer inputs, repeated: finds intermittent races.
    python scratch/dx_stress.py [reps]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ogc_b200 import backend
be = backend.get_backend(); lib = be.lib
_p = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
st = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
torch.manual_seed(0)
f32 = dict(dtype=torch.float32, device="cuda")


def run(fn, chain, B, N, M, cout, rows, synth, scatter, T):
    P = M * 64
    dz_prev = torch.zeros(B, rows, P, **f32); ab = torch.zeros(B, 4, 2, dtype=torch.float64, device="cuda")
    dg = torch.zeros(rows, **f32); db = torch.zeros(rows, **f32); dfeat = torch.zeros(B, N, rows, **f32)
    cs = torch.zeros(B, rows, 2, **f32)
    args = [B, N, M, 64, cout, T["W"].shape[1], 3 if scatter else 0, rows, None if synth else _p(T["dz"]), _p(T["go"]), cout, 0, _p(T["sel"]),
            _p(T["y"]), _p(T["coef"]), _p(T["W"])]
    if scatter:
        args += [None] * 8 + [_p(T["idx"]), _p(dfeat), rows, 0]
    else:
        args += [_p(T["yp"]), _p(T["ss"]), _p(T["mr"]), _p(T["gamma"]), _p(dz_prev), _p(ab), _p(dg), _p(db), None, None, 0, 0]
    if chain:
        args += [_p(cs), 0, 0]
    rc = fn(*args, st())
    assert rc == 0, rc
    torch.cuda.synchronize()
    return [dfeat] if scatter else [dz_prev, ab.float(), dg, db]


for (B, N, M, cout, rows, synth, scatter) in [(3, 1024, 512, 128, 128, False, False), (3, 1024, 512, 128, 128, True, False),
                                               (3, 1024, 512, 128, 128, False, True), (3, 2048, 1024, 64, 64, False, False),
                                               (3, 2048, 1024, 128, 64, True, False), (3, 2048, 1024, 64, 96, False, True)]:
    P = M * 64
    T = dict(y=torch.randn(B, cout, P, **f32), dz=torch.randn(B, cout, P, **f32), go=torch.randn(B, cout, M, **f32),
             sel=torch.randint(0, 64, (B, cout, M), dtype=torch.uint8, device="cuda"),
             coef=torch.randn(B, cout, 4, **f32) * 0.5, W=torch.randn(cout, rows + 3, **f32) * 0.1,
             yp=torch.randn(B, rows, P, **f32), ss=torch.randn(B, rows, 2, **f32), mr=torch.rand(B, 4, 2, **f32) + 0.5,
             gamma=torch.randn(rows, **f32), idx=torch.randint(0, N, (B, M, 64), dtype=torch.int32, device="cuda"))
    ref = run(lib.ogc_sa_mlp_layer_dx_tc, False, B, N, M, cout, rows, synth, scatter, T)
    bad = 0
    worst = 0.0
    for r in range(reps):
        out = run(lib.ogc_sa_chain_dx, True, B, N, M, cout, rows, synth, scatter, T)
        rel = max(float((a - b_).norm() / a.norm()) for a, b_ in zip(ref, out))
        worst = max(worst, rel)
        if rel > 1e-5:
            bad += 1
            if bad <= 3:
                d = (ref[0] - out[0]).abs()
                nz = (d > 1e-3 * ref[0].abs().max()).nonzero()
                print("   bad run", r, "rel", f"{rel:.2e}", "n bad", len(nz), "first", nz[:3].tolist(), "last", nz[-2:].tolist())
    print(f"cout {cout} rows {rows} synth {synth} scatter {scatter}: {bad}/{reps} bad runs, worst rel {worst:.2e}")
