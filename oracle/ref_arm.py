"""TEST / MEASUREMENT INFRASTRUCTURE -- the UNMODIFIED reference Python on a GPU, over the reference's own CUDA
extension (oracle/_ref/pointnet2_cuda_ref.so, built by oracle/ref_build.py from the sources under /root/reference).

This is the "reference pointnet2 CUDA ext" arm of BASELINE.json's >= 20x target:
`Trainer._train_it` of train_seg.py:47-86 itself, with `models/segnet_kitti.py`, `losses/seg_loss_unsup.py`,
`pointnet2/pointnet2.py`, torch.optim.Adam, LambdaLR and BNMomentumScheduler exactly as train_seg.py:248-351 wires
them -- none of this repository's models, losses, operators or kernels is imported here.

    python oracle/ref_arm.py stage                         # build container: copy the reference's .py files (and the
                                                           # yaml it is configured from) to oracle/_ref/py (git-ignored,
                                                           # shipped to the GPU box by gpurun like the .so)
    python oracle/ref_arm.py kittisf  --steps 10 --warmup 3 [--fp32] [--pairs 4] [--no-aug]
    python oracle/ref_arm.py oa_icp   --clouds 64 --icp-iter 20 --chunk 4
    python oracle/ref_arm.py flow     --npoint 2048 --batch 16 --iters 4 --steps 5

Each run prints ONE JSON line.  bench.py runs these in a SUBPROCESS, so the product process never maps
pointnet2_cuda_ref.so.  Nothing under /root/reference is read at run time.
"""
import argparse
import importlib.util
import json
import os
import shutil
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
PY_DIR = os.path.join(HERE, "_ref", "py")
SO_PATH = os.path.join(HERE, "_ref", "pointnet2_cuda_ref.so")
REF = "/root/reference"
STAGED = ["models", "losses", "utils", "metrics", "config", "pointnet2/pointnet2.py", "train_seg.py", "train_flow.py",
          "oa_icp.py"]


def stage():
    """Copy the reference's Python (no native sources) next to its built extension.  Build container only."""
    if not os.path.isdir(REF):
        return None
    if os.path.isdir(PY_DIR):
        shutil.rmtree(PY_DIR)
    os.makedirs(PY_DIR)
    for rel in STAGED:
        src, dst = os.path.join(REF, rel), os.path.join(PY_DIR, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if os.path.isdir(src):
            shutil.copytree(src, dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
        else:
            shutil.copy(src, dst)
    for d, _, files in os.walk(PY_DIR):
        os.chmod(d, 0o755)
        for f in files:
            os.chmod(os.path.join(d, f), 0o644)
    return PY_DIR


def available():
    return os.path.isdir(PY_DIR) and os.path.exists(SO_PATH)


def _load_by_path(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _setup_imports():
    """sys.path = the staged reference tree only (the repository root must NOT be importable here: its
    `pointnet2/` is a regular package and would shadow the reference's namespace package)."""
    import torch  # noqa: F401  (libtorch must be loaded before the extension)
    sys.path[:] = [p for p in sys.path if os.path.abspath(p or ".") not in (ROOT, HERE)]
    sys.path.insert(0, PY_DIR)
    ext = _load_by_path("pointnet2_cuda_ref", SO_PATH)
    sys.modules["pointnet2_cuda"] = ext                 # `import pointnet2_cuda as pointnet2` (pointnet2/pointnet2.py:7)
    # absent third-party modules the reference imports at module level but never touches on this path
    tbx = types.ModuleType("tensorboardX")

    class SummaryWriter:  # train_seg.py:44 creates one; nothing is logged by _train_it
        def __init__(self, *a, **k): pass
        def add_scalar(self, *a, **k): pass
        def flush(self): pass
    tbx.SummaryWriter = SummaryWriter
    sys.modules["tensorboardX"] = tbx
    mpl = types.ModuleType("matplotlib")
    mpl.pyplot = types.ModuleType("matplotlib.pyplot")
    sys.modules["matplotlib"] = mpl
    sys.modules["matplotlib.pyplot"] = mpl.pyplot
    for name in ("open3d", "pyquaternion", "skspatial", "skspatial.objects", "png"):
        sys.modules.setdefault(name, types.ModuleType(name))


def _data_module():
    # the synthetic generator is plain numpy/torch (no kernels); loaded by path so the repository root stays unimportable
    return _load_by_path("ogc_data_gen", os.path.join(ROOT, "ogc_b200", "data.py"))


def _fps_order_fn():
    import torch
    ext = sys.modules["pointnet2_cuda"]

    def fps(xyz, npoint):
        B, N, _ = xyz.shape
        out = torch.empty(B, npoint, dtype=torch.int32, device=xyz.device)
        temp = torch.full((B, N), 1e10, dtype=torch.float32, device=xyz.device)
        ext.furthest_point_sampling_wrapper(B, N, npoint, xyz.contiguous(), temp, out)
        return out
    return fps


def _set_precision(fp32):
    import torch
    torch.backends.cudnn.allow_tf32 = not fp32          # torch default: True -> the 1x1 convs run in TF32
    torch.backends.cuda.matmul.allow_tf32 = False       # torch default
    return "strict fp32 (cudnn.allow_tf32=False)" if fp32 else "torch defaults (cuDNN convs in TF32)"


def run_kittisf(args):
    """train_seg.py:248-351 (main) + :47-86 (_train_it), unmodified, on synthetic KITTI-SF-shaped batches."""
    import tempfile
    import numpy as np
    import torch
    import yaml
    _setup_imports()
    mode = _set_precision(args.fp32)
    import train_seg
    from torch import optim
    from torch.optim.lr_scheduler import LambdaLR
    cfg = yaml.load(open(os.path.join(PY_DIR, "config/seg/kittisf/kittisf_unsup.yaml")), Loader=yaml.FullLoader)
    a = argparse.Namespace(**cfg)
    a.batch_size = args.pairs
    train_seg.args = a                                   # lr_curve / bn_curve read the module-level `args`
    np.random.seed(a.random_seed)
    torch.manual_seed(a.random_seed)
    from models.segnet_kitti import MaskFormer3D
    segnet = MaskFormer3D(n_slot=a.segnet["n_slot"], n_point=a.segnet["n_point"], use_xyz=a.segnet["use_xyz"],
                          n_transformer_layer=a.segnet["n_transformer_layer"],
                          transformer_embed_dim=a.segnet["transformer_embed_dim"],
                          transformer_input_pos_enc=a.segnet["transformer_input_pos_enc"]).cuda()
    optimizer = optim.Adam(segnet.parameters(), lr=a.lr, weight_decay=a.weight_decay)
    lr_scheduler = LambdaLR(optimizer, lr_lambda=train_seg.lr_curve)
    bnm_scheduler = train_seg.BNMomentumScheduler(segnet, bn_lambda=train_seg.bn_curve)
    L = train_seg
    criterion = L.UnsupervisedOGCLoss(L.DynamicLoss(**a.loss["dynamic_loss_params"]),
                                      L.SmoothLoss(**a.loss["smooth_loss_params"]),
                                      L.InvarianceLoss(**a.loss["invariance_loss_params"]), L.EntropyLoss(), L.RankLoss(),
                                      weights=a.loss["weights"], start_steps=a.loss["start_steps"])
    trainer = L.Trainer(segnet=segnet, criterion=criterion, optimizer=optimizer,
                        aug_transform_epoch=a.aug_transform_epoch, ignore_npoint_thresh=a.ignore_npoint_thresh,
                        exp_base=tempfile.mkdtemp(prefix="ogc_ref_arm_"), lr_scheduler=lr_scheduler,
                        bnm_scheduler=bnm_scheduler)
    data = _data_module()
    aug = not args.no_aug
    dev = torch.device("cuda", 0)
    batches = [data.make_batch(500 + i, args.pairs, a.segnet["n_point"], aug=aug, fps_fn=_fps_order_fn(), device=dev)
               for i in range(2)]
    it0 = 100000                                         # past every start_step: all loss terms weighted in
    import warnings
    warnings.filterwarnings("ignore")
    for i in range(args.warmup):
        trainer._train_it(it0 + i, batches[i % 2], aug_transform=aug)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for i in range(args.steps):
        loss_dict, _, _ = trainer._train_it(it0 + i, batches[i % 2], aug_transform=aug)
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / args.steps
    clouds = args.pairs * (4 if aug else 2)
    return {"impl": "unmodified reference python", "config": "kittisf", "value": clouds / (ms * 1e-3), "unit": "clouds/s",
            "ms_per_step": ms, "steps": args.steps, "warmup": args.warmup, "precision": mode,
            "clouds_per_step": clouds, "loss": {k: float(v) for k, v in loss_dict.items()},
            "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30,
            "what": "train_seg.py Trainer._train_it + models/segnet_kitti.py + losses/seg_loss_unsup.py + "
                    "pointnet2/pointnet2.py over the reference's own CUDA extension (unchanged .cu, sm_100a), same B200"}


def run_oa_icp(args):
    """oa_icp.py:41-84 `object_aware_icp`, unmodified, run as chunks of `--chunk` clouds (the reference's own
    --test_batch_size 4 for KITTI-SF, README.md:238: B = 64 at once would need > 69 GB of N x N temporaries)."""
    import numpy as np
    import torch
    _setup_imports()
    _set_precision(args.fp32)
    import oa_icp
    data = _data_module()
    dev = torch.device("cuda", 0)
    N, K, B = 8192, 10, args.clouds
    batch = data.make_batch(900, B, N, aug=False, fps_fn=_fps_order_fn(), device=dev)
    pcs, segms, flows = batch[0].to(dev), batch[1].to(dev), batch[2].to(dev)
    # soft masks with the shape of segnet outputs: ground-truth segments (mod K) softened
    def soft(seg):
        onehot = torch.nn.functional.one_hot((seg.long() % K), K).float()
        return torch.softmax(onehot * 4.0 + 0.3 * torch.randn_like(onehot), -1)
    torch.manual_seed(3)
    m1, m2 = soft(segms[:, 0]), soft(segms[:, 1])
    pc1, pc2, flow = pcs[:, 0].contiguous(), pcs[:, 1].contiguous(), flows[:, 0].contiguous()

    def once():
        outs = []
        for c in range(0, B, args.chunk):
            sl = slice(c, c + args.chunk)
            with torch.no_grad():
                outs.append(oa_icp.object_aware_icp(pc1[sl], pc2[sl], flow[sl], m1[sl], m2[sl], icp_iter=args.icp_iter))
        return torch.cat(outs)
    for _ in range(args.warmup):
        once()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(args.steps):
        once()
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / args.steps
    return {"impl": "unmodified reference python", "config": "oa_icp", "value": B / (ms * 1e-3), "unit": "clouds/s",
            "ms_per_step": ms, "steps": args.steps, "warmup": args.warmup, "clouds_per_step": B, "chunk": args.chunk,
            "icp_iter": args.icp_iter, "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30,
            "what": "oa_icp.object_aware_icp (unmodified) in chunks of %d clouds over the reference extension" % args.chunk}


def run_flow(args):
    """train_flow.py:225-283 (main) + :59-92 (_train_it), unmodified: models/flownet_ogcdr.FlowStep3D forward (iters) +
    losses/flow_loss_unsup + backward + Adam, configured from config/flow/ogcdr/ogcdr_unsup.yaml."""
    import tempfile
    import numpy as np
    import torch
    import yaml
    _setup_imports()
    mode = _set_precision(args.fp32)
    import train_flow
    from torch import optim
    from torch.optim.lr_scheduler import LambdaLR
    cfg = yaml.load(open(os.path.join(PY_DIR, "config/flow/ogcdr/ogcdr_unsup.yaml")), Loader=yaml.FullLoader)
    a = argparse.Namespace(**cfg)
    a.batch_size = args.batch
    a.flownet = dict(a.flownet, npoint=args.npoint)
    a.model_iters = args.iters
    a.loss = dict(a.loss, iters_w=([0.5] + [0.3] * (args.iters - 1)))
    train_flow.args = a
    np.random.seed(a.random_seed)
    torch.manual_seed(a.random_seed)
    from models.flownet_ogcdr import FlowStep3D
    T = train_flow
    flownet = FlowStep3D(npoint=a.flownet["npoint"], use_instance_norm=a.flownet["use_instance_norm"],
                         loc_flow_nn=a.flownet["loc_flow_nn"], loc_flow_rad=a.flownet["loc_flow_rad"],
                         k_decay_fact=a.flownet["k_decay_fact"]).cuda()
    optimizer = optim.Adam(flownet.parameters(), lr=a.lr, weight_decay=a.weight_decay)
    lr_scheduler = LambdaLR(optimizer, lr_lambda=T.lr_curve)
    bnm_scheduler = T.BNMomentumScheduler(flownet, bn_lambda=T.bn_curve)
    criterion = T.UnsupervisedFlowStep3DLoss(chamfer_loss=T.ChamferLoss(**a.loss["chamfer_loss_params"]),
                                             smooth_loss=T.SmoothLoss(**a.loss["smooth_loss_params"]),
                                             iters_w=a.loss["iters_w"], weights=a.loss["weights"])
    trainer = T.Trainer(flownet=flownet, model_iters=a.model_iters, criterion=criterion, optimizer=optimizer,
                        exp_base=tempfile.mkdtemp(prefix="ogc_ref_arm_"), lr_scheduler=lr_scheduler,
                        bnm_scheduler=bnm_scheduler)
    batches = [flow_batch(31 + i, args.batch, args.npoint) for i in range(2)]
    import warnings
    warnings.filterwarnings("ignore")
    for i in range(args.warmup):
        trainer._train_it(i, batches[i % 2])
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for i in range(args.steps):
        loss_dict = trainer._train_it(100 + i, batches[i % 2])
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / args.steps
    return {"impl": "unmodified reference python", "config": "ogcdr_flow", "value": args.batch / (ms * 1e-3),
            "unit": "pairs/s", "ms_per_step": ms, "steps": args.steps, "warmup": args.warmup, "precision": mode,
            "npoint": args.npoint, "batch": args.batch, "iters": args.iters,
            "loss": {k: float(v) for k, v in loss_dict.items()},
            "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30,
            "what": "train_flow.py Trainer._train_it + models/flownet_ogcdr.FlowStep3D + losses/flow_loss_unsup "
                    "(unmodified) over the reference's own CUDA extension, same B200"}


def flow_batch(seed, batch, npoint):
    """Synthetic OGC-DR-shaped batch: 4-8 rigid parts in [-0.5,0.5]^3 moving by <= 0.1;
    layout of datasets/dataset_ogcdr.py: pcs (b,2,N,3), segms (b,2,N), flows (b,2,N,3), valids."""
    import numpy as np
    import torch
    rng = np.random.default_rng(seed)
    P, S, F_ = [], [], []
    for _ in range(batch):
        n_obj = int(rng.integers(4, 9))
        seg = rng.integers(0, n_obj, npoint)
        centre = rng.uniform(-0.35, 0.35, (n_obj, 3))
        size = rng.uniform(0.05, 0.15, (n_obj, 3))
        pc1 = centre[seg] + rng.uniform(-1, 1, (npoint, 3)) * size[seg]
        ang = rng.uniform(-0.2, 0.2, n_obj)
        t = rng.uniform(-0.05, 0.05, (n_obj, 3))
        c, s_ = np.cos(ang)[seg], np.sin(ang)[seg]
        rel = pc1 - centre[seg]
        moved = np.stack([c * rel[:, 0] + s_ * rel[:, 2], rel[:, 1], -s_ * rel[:, 0] + c * rel[:, 2]], 1) + centre[seg] + t[seg]
        flow1 = moved - pc1
        perm = rng.permutation(npoint)
        pc2 = (moved + rng.normal(0, 0.002, moved.shape))[perm]
        P.append(np.stack([pc1, pc2])); S.append(np.stack([seg, seg[perm]])); F_.append(np.stack([flow1, -flow1[perm]]))
    pcs = torch.from_numpy(np.stack(P).astype(np.float32))
    segms = torch.from_numpy(np.stack(S).astype(np.int32))
    flows = torch.from_numpy(np.stack(F_).astype(np.float32))
    return pcs, segms, flows, torch.ones(segms.shape)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("what", choices=["stage", "kittisf", "oa_icp", "flow"])
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--fp32", action="store_true")
    ap.add_argument("--pairs", type=int, default=4)
    ap.add_argument("--no-aug", action="store_true")
    ap.add_argument("--clouds", type=int, default=64)
    ap.add_argument("--icp-iter", type=int, default=20)
    ap.add_argument("--chunk", type=int, default=4)
    ap.add_argument("--npoint", type=int, default=2048)
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--iters", type=int, default=4)
    args = ap.parse_args()
    if args.what == "stage":
        print(stage())
        return
    if not available():
        print(json.dumps({"impl": "unmodified reference python", "unavailable": "oracle/_ref/py or the extension is missing"}))
        return
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    out = {"kittisf": run_kittisf, "oa_icp": run_oa_icp, "flow": run_flow}[args.what](args)
    sys.stdout.flush()
    os.write(real_stdout, (json.dumps(out) + "\n").encode())


if __name__ == "__main__":
    main()
