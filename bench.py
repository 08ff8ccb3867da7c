#!/usr/bin/env python
"""bench.py -- the headline benchmark of BASELINE.json:

    point-clouds/sec for (segnet_kitti forward + UnsupervisedOGCLoss + backward + Adam) at 8192 points

    python bench.py --gpus N --steps K --warmup W            # this repo (sm_100a kernels), N ranks via torchrun
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm: reference-structure step on the host cores

One "step" = one pass of the hot path over one batch: `Trainer._train_it` (train_seg.py:47-86) restated in
ogc_b200/train.py.  Workload (configs[1] at the reference's own training settings,
config/seg/kittisf/kittisf_unsup.yaml): b = 4 KITTI-SF-like pairs per GPU, augmentation on (t = 4 views) =>
16 clouds x 8192 points per GPU per step, n_slot 10, all three loss terms active.  Weak scaling: per-GPU
work is fixed; ranks hold different seeded shards and exchange one NCCL all-reduce (grads + NaN counter).

Prints ONE JSON line (rank 0).  `value` = clouds/s with the batch already resident in HBM; `e2e` = the
same through the public step API with pinned-host inputs copied in and the loss dict read back every step.
`roofline` is for the dominant libogc_b200 kernel family of the step (CUDA events on the launching stream, side-stream
overlap switched off for that pass; `traffic` = DRAM bytes per span from the committed ncu capture, profiles/ncu_traffic.json);
`ops` lists every kernel family of ours (FPS and ball_query included, as BASELINE.json's metric asks).
`cpu_baseline` (rank 0, N=1) times the CPU port (oracle kernels + the same torch step on the host cores)
on a bounded sample (10 steps of one pair); `ref_cuda_ext` (N=1) times the reference's own CUDA extension under the
reference's op sequence on the same GPU.  stdout carries exactly the JSON line (library banners go to stderr).
With N > 1 the NCCL all-reduce, Adam and the log read-back follow the step graph eagerly.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

N_POINT = 8192
N_SLOT = 10
METRIC = "point-clouds/sec (8192 pts, segnet fwd+OGC-loss bwd)"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("hbm_gbs") is not None
    return 6650.0, False


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def build_trainer(device, world, variant_impl="b200"):
    from ogc_b200.segnet import MaskFormer3D
    from ogc_b200.losses import build_ogc_loss, KITTISF_LOSS_CFG
    from ogc_b200.train import SegTrainer
    torch.manual_seed(10)                       # config/seg/kittisf/kittisf_unsup.yaml:3 random_seed
    net = MaskFormer3D(n_slot=N_SLOT, n_point=N_POINT, variant="kitti").to(device)
    crit = build_ogc_loss(KITTISF_LOSS_CFG)
    return SegTrainer(net, crit, lr=1e-3, global_batch_size=4 * world, world_size=world)


def run_cpu_port(steps, warmup, pairs_per_step, threads):
    """Reference-structure step on the host cores: the same torch step + the CPU oracle kernels
    (the reference itself has no CPU path: SURVEY.md 8c).  Returns (clouds/s, ms/step, clouds/step)."""
    from ogc_b200 import backend, data
    from oracle.pointnet2_oracle import OracleBackend
    torch.set_num_threads(threads)
    os.environ.setdefault("OMP_NUM_THREADS", str(threads))
    prev = backend.set_backend(OracleBackend())
    try:
        trainer = build_trainer(torch.device("cpu"), 1)
        batch = data.make_batch(1234, pairs_per_step, N_POINT, aug=False)
        for i in range(warmup):
            trainer.train_step(2000 + i, batch, aug_transform=False)
        t0 = time.perf_counter()
        for i in range(steps):
            trainer.train_step(3000 + i, batch, aug_transform=False)
        dt = (time.perf_counter() - t0) / max(steps, 1)
    finally:
        backend.set_backend(prev)
    clouds = pairs_per_step * 2
    return clouds / dt, dt * 1e3, clouds


def run_ref_cuda_ext(device, steps, warmup, pairs, aug):
    """The reference's OWN CUDA extension (oracle/_ref, built from /root/reference) under the reference's op
    sequence (segnet.REFERENCE_FAITHFUL / losses.REFERENCE_FAITHFUL) on this GPU: the denominator of
    BASELINE.json's ">= 20x the reference pointnet2 CUDA ext" target.  Returns None when the extension was not
    built.  The reference Python itself cannot travel to the GPU box; the mirror reproduces its call sequence."""
    from oracle import refext
    if not refext.available():
        return None
    from ogc_b200 import backend, data, losses, segnet
    prev = backend.set_backend(refext.RefExtBackend())
    segnet.REFERENCE_FAITHFUL = losses.REFERENCE_FAITHFUL = True
    try:
        trainer = build_trainer(device, 1)
        batches = [tuple(x.to(device) for x in data.make_batch(500 + i, pairs, N_POINT, aug=aug)) for i in range(2)]
        for i in range(warmup):
            trainer.train_step(100000 + i, batches[i % 2], aug_transform=aug)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for i in range(steps):
            trainer.train_step(100000 + i, batches[i % 2], aug_transform=aug)
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e) / steps
    finally:
        segnet.REFERENCE_FAITHFUL = losses.REFERENCE_FAITHFUL = False
        backend.set_backend(prev)
        torch.cuda.empty_cache()
    clouds = pairs * (4 if aug else 2)
    return {"value": clouds / (ms * 1e-3), "unit": "clouds/s", "ms_per_step": ms, "steps": steps,
            "what": "reference pointnet2 CUDA extension (unchanged .cu, sm_100a) + reference op sequence "
                    "(cuDNN TF32 convs, diag_embed Kabsch, per-scale kNN, SVD nuclear norm) on the same B200"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--pairs", type=int, default=4, help="KITTI-SF pairs per GPU per step (reference batch_size)")
    ap.add_argument("--no-aug", action="store_true")
    ap.add_argument("--eager", action="store_true", help="one launch per kernel instead of CUDA-graph replay")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ref-ext", action="store_true", help="skip timing the reference CUDA extension arm")
    args = ap.parse_args()

    # stdout carries exactly ONE line, the JSON: everything else any library prints there (NCCL's version banner ...)
    # is redirected to stderr at the file-descriptor level
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(obj) + "\n").encode())

    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    aug = not args.no_aug
    t_views = 4 if aug else 2
    cores = os.cpu_count() or 1
    workload = (f"kittisf_unsup step: {args.pairs} pairs/GPU x {t_views} views = {args.pairs * t_views} clouds/GPU/step, "
                f"{N_POINT} pts, n_slot {N_SLOT}, segnet_kitti + OGC loss (dynamic+smooth+invariance) + bwd + Adam")

    # ------------------------------------------------------------------ CPU arm ("reference")
    if args.impl == "reference":
        if rank != 0:
            return
        value, ms, clouds = run_cpu_port(args.steps, min(args.warmup, 1), 1, cores)
        line = {"metric": METRIC, "value": value, "unit": "clouds/s", "n_gpus": args.gpus, "steps": args.steps,
                "warmup": min(args.warmup, 1), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
                "config": {"workload": workload, "sample": "1 pair (2 clouds), no augmentation, per step"},
                "cpu_baseline": {"value": value, "unit": "clouds/s", "cores": cores, "kind": "port",
                                 "sample": f"{args.steps} steps x 1 pair (2 clouds x {N_POINT} pts), fwd+loss+bwd+Adam"},
                "e2e": {"value": value, "unit": "clouds/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        emit(line)
        return

    # ------------------------------------------------------------------ B200 arm
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=device)
    from ogc_b200 import backend, data
    be = backend.get_backend()
    trainer = build_trainer(device, world)

    n_batches = 4
    batches = [data.make_batch(1000 * rank + i, args.pairs, N_POINT, aug=aug, fps_fn=be.fps, device=device)
               for i in range(n_batches)]
    resident = [tuple(x.to(device) for x in b) for b in batches]
    h2d = sum(x.numel() * x.element_size() for x in (batches[0][0], batches[0][2]))
    clouds_per_step = args.pairs * t_views * world
    it0 = 100000        # past every start_step: all loss terms weighted in

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(src, steps):
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = be.launches
        s.record()
        last = None
        for i in range(steps):
            last = step_fn(it0 + i, src[i % n_batches], aug_transform=aug)
        e.record()
        barrier()
        ms = s.elapsed_time(e)
        if world > 1:
            t = torch.tensor([ms], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / steps, (be.launches - l0) / steps, last

    # the whole step is one CUDA-graph launch (device-side Hungarian / NaN guard / Adam state); --eager keeps the
    # one-launch-per-kernel path
    step_fn = trainer.train_step if args.eager else trainer.train_step_graphed
    for i in range(max(args.warmup, 3)):
        step_fn(it0 + i, resident[i % n_batches], aug_transform=aug)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_step, launches, _ = timed(resident, args.steps)
    ms_e2e, _, last_dict = timed(batches, args.steps)            # pinned host -> device every step, loss dict back
    clocks = sampler.stop() if rank == 0 else None

    # per-kernel timing of OUR kernels over two more steps (CUDA events on the launching stream)
    # (the FPS side-stream overlap is switched off here so that every kernel is timed alone)
    backend.TIMER.enabled = True
    backend.TIMER.reset()
    trainer.overlap_geometry = False
    for i in range(2):
        trainer.train_step(it0 + i, resident[i % n_batches], aug_transform=aug)
    torch.cuda.synchronize()
    trainer.overlap_geometry = True
    backend.TIMER.enabled = False
    ops = backend.TIMER.summary()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, measured = measured_peaks()
    op_rows = {}
    for name, d in ops.items():
        per_launch_ms = d["ms"] / d["calls"]
        gbs = d["bytes"] / d["calls"] / (per_launch_ms * 1e-3) / 1e9
        op_rows[name] = {"calls_per_step": d["calls"] / 2, "ms_per_step": d["ms"] / 2, "avg_ms": per_launch_ms,
                         "alg_gbs": gbs, "frac_of_hbm_peak": gbs / peak}
    top = max(op_rows, key=lambda k: op_rows[k]["ms_per_step"]) if op_rows else None
    roofline = None
    if top:
        # DRAM bytes per launch of that kernel family from the committed `ncu --set full` capture (profiles/)
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get(top, {}).get("dram_bytes_per_launch")
        roofline = {"kernel": top, "bound": "hbm", "achieved": op_rows[top]["alg_gbs"], "peak": peak, "unit": "GB/s",
                    "frac": op_rows[top]["alg_gbs"] / peak, "traffic": traffic,
                    "alg_bytes_per_launch": ops[top]["bytes"] / ops[top]["calls"],
                    "peak_source": "MEASURED_PEAKS.json (measured)" if measured else "fallback 6650 GB/s",
                    "share_of_step": op_rows[top]["ms_per_step"] / ms_step,
                    "note": "algorithmic bytes per launch / CUDA-event duration; see DESIGN.md for the byte formulas"}

    line = {"metric": METRIC, "value": clouds_per_step / (ms_step * 1e-3), "unit": "clouds/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload, "parallelism": f"dp{world}",
                       "launch": "eager" if args.eager else "one CUDA graph per step",
                       "l2": "per-step working set (GBs of activations) >> 126 MB L2; inputs cycle over 4 distinct batches"},
            "e2e": {"value": clouds_per_step / (ms_e2e * 1e-3), "unit": "clouds/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4 * len(last_dict)},
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "ops": op_rows,
            "loss": last_dict}

    if world == 1 and not args.no_ref_ext:
        ref = run_ref_cuda_ext(device, 2, 1, args.pairs, aug)
        if ref is not None:
            ref["speedup_e2e"] = line["e2e"]["value"] / ref["value"]
            line["ref_cuda_ext"] = ref
    if world == 1 and not args.no_cpu_baseline:
        cpu_steps = 10                      # ~10-15 s of CPU work on the box's host cores
        v, ms, clouds = run_cpu_port(cpu_steps, 1, 1, cores)
        line["cpu_baseline"] = {"value": v, "unit": "clouds/s", "cores": cores, "kind": "port", "ms_per_step": ms,
                                "sample": f"{cpu_steps} steps (+1 warm-up) x 1 pair ({clouds} clouds x {N_POINT} pts, no aug): "
                                          "oracle kernels + the same torch step on the host cores"}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
