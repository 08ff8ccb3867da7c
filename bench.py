#!/usr/bin/env python
"""bench.py -- the benchmarks of BASELINE.json, one JSON line per run.

    python bench.py --gpus N --steps K --warmup W                      # headline: configs[1] at the reference's training
                                                                       # settings (kittisf_unsup.yaml: 4 pairs x 4 views)
    python bench.py --config train8|strong32|oa_icp|ogcdr_flow|ogcdr_flow4096|sapien_cpu ...
    python bench.py --impl reference [--config ...] --gpus N --steps K --warmup W     # CPU arm (oracle port on the host cores)

Headline metric: point-clouds/sec for (segnet_kitti forward + UnsupervisedOGCLoss + backward + Adam) at 8192 points.
One "step" = one pass of `Trainer._train_it` (train_seg.py:47-86, restated in ogc_b200/train.py) over one batch:
b = 4 KITTI-SF-like pairs per GPU with augmentation (t = 4 views) => 16 clouds x 8192 points per GPU per step,
n_slot 10, all three loss terms active.  Weak scaling: per-GPU work fixed, ranks hold different seeded shards and
exchange one NCCL all-reduce (grads + NaN counter).

    value        clouds/s with the batch already resident in HBM (CUDA events, max over ranks)
    e2e          the same through the public step API with pinned-host inputs copied in and the loss dict read back
    roofline     the dominant libogc_b200 kernel family of the step's critical path (the single-wave kernels hidden on side
                 streams -- FPS chain, three_nn, Hungarian, nuclear norm -- are in `ops` only).  A contraction kernel is
                 placed by its arithmetic intensity (3xTF32 work per algorithmic byte) against the ridge of the two
                 measured peaks: "tensor" (useful fp32-equivalent FLOP/s against the TF32 peak = half the measured bf16
                 cuBLAS peak) or "hbm" (SURVEY 8d algorithmic bytes / time against the measured copy bandwidth; the
                 tensor fractions alongside); `traffic` = DRAM bytes per launch from the committed ncu capture; `step` =
                 the whole-step HBM fraction of SURVEY 8d
    The step loop announces the NEXT batch to the trainer (a loader one batch ahead): its clouds are copied in and their
    first-level FPS centres sampled under the current step -- one copy and one FPS per step either way (--no-prefetch off).
    ops          every kernel family of ours (FPS and ball_query GB/s included, as BASELINE.json's metric asks)
    cpu_baseline the CPU port (oracle kernels + the same torch step, all host cores) on a bounded sample   } run in
    ref_cuda_ext the UNMODIFIED reference Python over the reference's own CUDA extension on the same GPU,  } SUBPROCESSES:
                 torch-default (TF32 convs) and strict-fp32 modes, >= 10 timed steps (oracle/ref_arm.py)    } the product
                 -- the denominator of BASELINE.json's ">= 20x" target                                      } process maps
                                                                                                            } libogc_b200 only
Other configs (`--config`): train8 = configs[3] (8 pairs/GPU); strong32 = 32 clouds split over the ranks (strong
scaling); oa_icp = configs[4] (64 clouds, icp_iter 20; the reference runs chunks of 4); ogcdr_flow[4096] = configs[2]
(FlowStep3D b=16 iters=4 + flow loss + backward + Adam at 2048 / 4096 points; `ops` / `roofline` from two eager steps,
the 3xTF32-forward variant as a labelled extra); sapien_cpu = configs[0] (CPU plumbing).
stdout carries exactly the JSON line (library banners go to stderr).
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
ALG_BYTES_PER_CLOUD = 42.5e6          # perfectly fused units, fwd + loss + bwd


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d.get("hbm_gbs", 6650.0), "bf16_tflops": d.get("bf16_tflops", 1590.0), "measured": True}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "measured": False}


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


def build_trainer(device, world, pairs):
    from ogc_b200.segnet import MaskFormer3D
    from ogc_b200.losses import build_ogc_loss, KITTISF_LOSS_CFG
    from ogc_b200.train import SegTrainer
    torch.manual_seed(10)                       # config/seg/kittisf/kittisf_unsup.yaml:3 random_seed
    net = MaskFormer3D(n_slot=N_SLOT, n_point=N_POINT, variant="kitti").to(device)
    crit = build_ogc_loss(KITTISF_LOSS_CFG)
    return SegTrainer(net, crit, lr=1e-3, global_batch_size=pairs * world, world_size=world)


# ---------------------------------------------------------------------------------------------------------------------
# subprocess legs (checkers / baselines never share the product process)
# ---------------------------------------------------------------------------------------------------------------------
def _run_json(cmd, timeout):
    """Run a helper that prints one JSON line on stdout; {"unavailable": why} on any failure."""
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    except subprocess.TimeoutExpired:
        return {"unavailable": f"timeout after {timeout}s: {' '.join(cmd[1:4])}"}
    for ln in reversed(r.stdout.strip().splitlines()):
        if ln.startswith("{"):
            try:
                return json.loads(ln)
            except json.JSONDecodeError:
                pass
    return {"unavailable": f"rc={r.returncode}: {r.stderr.strip().splitlines()[-1][:200] if r.stderr.strip() else 'no output'}"}


def ref_arm(*args, timeout=600):
    """oracle/ref_arm.py: the UNMODIFIED reference Python over the reference's own CUDA extension, in a subprocess."""
    return _run_json([sys.executable, os.path.join(ROOT, "oracle", "ref_arm.py"), *map(str, args)], timeout)


def cpu_arm(config, steps, warmup, timeout=900):
    """This file's --impl reference arm (CPU port), in a subprocess."""
    return _run_json([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--config", config,
                      "--steps", str(steps), "--warmup", str(warmup)], timeout)


# ---------------------------------------------------------------------------------------------------------------------
# CPU arm ("reference"): the oracle port on the host cores.  The only place besides tests/ and smoke() that executes oracle/.
# ---------------------------------------------------------------------------------------------------------------------
def run_cpu_seg(steps, warmup, pairs, aug, threads):
    """Reference-structure step on the host cores: the same torch step + the CPU oracle kernels
    (the reference itself has no CPU path: SURVEY.md 8c).  Returns (clouds/s, ms/step, clouds/step)."""
    from ogc_b200 import backend, data
    from oracle.pointnet2_oracle import OracleBackend
    torch.set_num_threads(threads)
    os.environ.setdefault("OMP_NUM_THREADS", str(threads))
    prev = backend.set_backend(OracleBackend())
    try:
        trainer = build_trainer(torch.device("cpu"), 1, pairs)
        batch = data.make_batch(1234, pairs, N_POINT, aug=aug)
        for i in range(warmup):
            trainer.train_step(100000 + i, batch, aug_transform=aug)
        t0 = time.perf_counter()
        for i in range(steps):
            trainer.train_step(100100 + i, batch, aug_transform=aug)
        dt = (time.perf_counter() - t0) / max(steps, 1)
    finally:
        backend.set_backend(prev)
    clouds = pairs * (4 if aug else 2)
    return clouds / dt, dt * 1e3, clouds


def run_cpu_sapien(steps, warmup, threads):
    """configs[0]: one SAPIEN 512-point cloud through segnet_sapien (forward) on CPU over the oracle kernels."""
    from ogc_b200 import backend
    from ogc_b200.segnet import MaskFormer3D
    from oracle.pointnet2_oracle import OracleBackend
    import numpy as np
    torch.set_num_threads(threads)
    prev = backend.set_backend(OracleBackend())
    try:
        torch.manual_seed(10)
        net = MaskFormer3D(n_slot=8, n_point=512, variant="sapien")
        net.eval()
        pc = torch.from_numpy(np.random.default_rng(10).uniform(-0.5, 0.5, (1, 512, 3)).astype(np.float32))
        with torch.no_grad():
            for _ in range(warmup):
                net(pc, pc)
            t0 = time.perf_counter()
            for _ in range(steps):
                mask = net(pc, pc)
            dt = (time.perf_counter() - t0) / max(steps, 1)
    finally:
        backend.set_backend(prev)
    assert mask.shape == (1, 512, 8)
    return 1.0 / dt, dt * 1e3


def run_cpu_icp(clouds, icp_iter, threads):
    from ogc_b200 import backend, icp
    from oracle.pointnet2_oracle import OracleBackend
    torch.set_num_threads(threads)
    prev = backend.set_backend(OracleBackend())
    try:
        pc1, pc2, flow, m1, m2 = icp_inputs(clouds, torch.device("cpu"))
        t0 = time.perf_counter()
        with torch.no_grad():
            icp.object_aware_icp(pc1, pc2, flow, m1, m2, icp_iter=icp_iter)
        dt = time.perf_counter() - t0
    finally:
        backend.set_backend(prev)
    return clouds / dt, dt * 1e3


def run_cpu_flow(npoint, batch, iters, steps, threads):
    from ogc_b200 import backend
    from oracle.pointnet2_oracle import OracleBackend
    torch.set_num_threads(threads)
    prev = backend.set_backend(OracleBackend())
    try:
        step, _ = flow_step_fn(npoint, batch, iters, torch.device("cpu"))
        step(0)
        t0 = time.perf_counter()
        for i in range(steps):
            step(1 + i)
        dt = (time.perf_counter() - t0) / max(steps, 1)
    finally:
        backend.set_backend(prev)
    return batch / dt, dt * 1e3


# ---------------------------------------------------------------------------------------------------------------------
# workloads shared by the arms
# ---------------------------------------------------------------------------------------------------------------------
def icp_inputs(B, device, N=N_POINT, K=N_SLOT):
    """KITTI-SF-like pairs + soft masks shaped like segnet outputs (same generator as oracle/ref_arm.py:run_oa_icp)."""
    from ogc_b200 import data
    batch = data.make_batch(900, B, N, aug=False)
    pcs, segms, flows = batch[0].to(device), batch[1].to(device), batch[2].to(device)
    torch.manual_seed(3)

    def soft(seg):
        onehot = torch.nn.functional.one_hot((seg.long() % K), K).float()
        return torch.softmax(onehot * 4.0 + 0.3 * torch.randn_like(onehot), -1)
    m1, m2 = soft(segms[:, 0]), soft(segms[:, 1])
    return pcs[:, 0].contiguous(), pcs[:, 1].contiguous(), flows[:, 0].contiguous(), m1.contiguous(), m2.contiguous()


def flow_step_fn(npoint, batch, iters, device):
    """FlowStep3D forward (iters) + unsupervised flow loss + backward + NaN guard + Adam on OGC-DR-shaped synthetic pairs
    (train_flow.py:59-92 with config/flow/ogcdr/ogcdr_unsup.yaml; generator shared with oracle/ref_arm.py).
    On the GPU the step is ogc_b200.train.FlowTrainer's single CUDA-graph launch."""
    import importlib.util
    from ogc_b200.flownet import FlowStep3D, build_flow_loss, OGCDR_FLOW_LOSS_CFG
    from ogc_b200.train import FlowTrainer
    spec = importlib.util.spec_from_file_location("ogc_ref_arm_gen", os.path.join(ROOT, "oracle", "ref_arm.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)                       # only its pure-numpy batch generator is used here
    torch.manual_seed(10)
    net = FlowStep3D(npoint=npoint, use_instance_norm=False, loc_flow_nn=8, loc_flow_rad=0.05).to(device)
    crit = build_flow_loss(dict(OGCDR_FLOW_LOSS_CFG, iters_w=[0.5] + [0.3] * (iters - 1)))
    trainer = FlowTrainer(net, crit, iters, lr=1e-3)
    batches = [gen.flow_batch(31 + i, batch, npoint) for i in range(2)]
    if device.type == "cuda":
        batches = [tuple(x.pin_memory() for x in b) for b in batches]
    graphed = device.type == "cuda" and os.environ.get("OGC_FLOW_EAGER", "0") != "1"

    resident = [tuple(x.to(device) for x in b) for b in batches]

    def step(i, host_inputs=True):
        b = (batches if host_inputs else resident)[i % 2]
        return trainer.train_step_graphed(b) if graphed else trainer.train_step(b)
    h2d = batches[0][0].numel() * 4
    step.trainer, step.resident = trainer, resident
    return step, h2d


def summarise_ops(ops, nsteps, peaks):
    """backend.TIMER.summary() -> per-family rows: calls / ms per step, algorithmic GB/s against the measured HBM peak,
    useful TFLOP/s; a tcgen05 contraction family is placed by its arithmetic intensity (3xTF32 work per algorithmic byte)
    against the ridge of the two measured peaks."""
    hbm_peak, tf32_peak = peaks["hbm_gbs"], peaks["bf16_tflops"] / 2.0
    op_rows = {}
    for name, d in ops.items():
        per_launch_ms = d["ms"] / d["calls"]
        row = {"calls_per_step": d["calls"] / nsteps, "ms_per_step": d["ms"] / nsteps, "avg_ms": per_launch_ms,
               "alg_bytes_per_launch": d["bytes"] / d["calls"]}
        if d.get("flops"):
            tf = d["flops"] / d["calls"] / (per_launch_ms * 1e-3) / 1e12
            row["useful_tflops"] = tf
            if "_tc" in name or "chain" in name or "_tma" in name:      # tcgen05 kernels; the narrow / SIMT ones are fp32 FFMA
                # roofline of a contraction kernel: by arithmetic intensity of the ISSUED tensor work (3 passes of the TF32
                # split) against the ridge of the two measured peaks
                intensity = 3.0 * d["flops"] / max(d["bytes"], 1)
                ridge = tf32_peak * 1e12 / (hbm_peak * 1e9)
                row.update({"bound": "tensor" if intensity > ridge else "hbm", "flop_per_byte_3xtf32": intensity,
                            "ridge_flop_per_byte": ridge, "frac_of_tf32_peak": tf / tf32_peak,
                            "frac_of_tf32_peak_3xtf32_work": 3 * tf / tf32_peak})
        gbs = d["bytes"] / d["calls"] / (per_launch_ms * 1e-3) / 1e9
        row.update({"alg_gbs": gbs, "frac_of_hbm_peak": gbs / hbm_peak})
        op_rows[name] = row
    return op_rows


def time_cuda(fn, steps, warmup):
    for i in range(warmup):
        fn(i)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    out = None
    for i in range(steps):
        out = fn(warmup + i)
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / steps, out


def emit_factory():
    # stdout carries exactly ONE line, the JSON: everything else any library prints there (NCCL's version banner ...)
    # is redirected to stderr at the file-descriptor level
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(obj) + "\n").encode())
    return emit


SEG_CONFIGS = {"kittisf": 4, "train8": 8, "strong32": None}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="kittisf",
                    choices=["kittisf", "train8", "strong32", "oa_icp", "ogcdr_flow", "ogcdr_flow4096", "sapien_cpu"])
    ap.add_argument("--pairs", type=int, default=None, help="KITTI-SF pairs per GPU per step (default: the config's)")
    ap.add_argument("--no-aug", action="store_true")
    ap.add_argument("--eager", action="store_true", help="one launch per kernel instead of CUDA-graph replay")
    ap.add_argument("--no-prefetch", action="store_true",
                    help="do not announce the next batch to the trainer (first-level FPS inside the step instead of under the previous one)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ref-ext", action="store_true", help="skip timing the reference CUDA extension arm")
    args = ap.parse_args()
    emit = emit_factory()

    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    aug = not args.no_aug
    t_views = 4 if aug else 2
    cores = os.cpu_count() or 1
    cfg = args.config
    if cfg in SEG_CONFIGS:
        pairs = args.pairs or SEG_CONFIGS[cfg] or max(1, 8 // world)       # strong32: 32 clouds over the ranks
    else:
        pairs = 0
    seg_workload = (f"kittisf_unsup step: {pairs} pairs/GPU x {t_views} views = {pairs * t_views} clouds/GPU/step, "
                    f"{N_POINT} pts, n_slot {N_SLOT}, segnet_kitti + OGC loss (dynamic+smooth+invariance) + bwd + Adam")

    # ------------------------------------------------------------------ CPU arm ("reference")
    if args.impl == "reference":
        if rank != 0:
            return
        w = min(args.warmup, 1)
        base = {"n_gpus": args.gpus, "steps": args.steps, "warmup": w, "higher_is_better": True, "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "impl": "reference", "gpu_launches": 0}
        if cfg in SEG_CONFIGS:
            value, ms, clouds = run_cpu_seg(args.steps, w, pairs, aug, cores)
            line = dict(base, metric=METRIC, value=value, unit="clouds/s", ms_per_step=ms,
                        scaling="strong" if cfg == "strong32" else "weak",
                        config={"workload": seg_workload, "sample": f"the same {clouds}-cloud step, {args.steps} timed steps"},
                        cpu_baseline={"value": value, "unit": "clouds/s", "cores": cores, "kind": "port", "ms_per_step": ms,
                                      "sample": f"{args.steps} steps (+{w} warm-up) of the full {clouds}-cloud step: oracle "
                                                "kernels + the same torch step on the host cores"})
        elif cfg == "sapien_cpu":
            value, ms = run_cpu_sapien(args.steps, w, cores)
            line = dict(base, metric="point-clouds/sec (512 pts, segnet_sapien forward, CPU plumbing)", value=value,
                        unit="clouds/s", ms_per_step=ms, scaling="weak",
                        config={"workload": "configs[0]: one SAPIEN 512-pt cloud, segnet_sapien forward on CPU over the oracle kernels"},
                        cpu_baseline={"value": value, "unit": "clouds/s", "cores": cores, "kind": "port", "sample": f"{args.steps} forwards"})
        elif cfg == "oa_icp":
            value, ms = run_cpu_icp(4, 2, cores)
            line = dict(base, metric="point-clouds/sec (8192 pts, object-aware ICP)", value=value * 2 / 20, unit="clouds/s",
                        ms_per_step=ms, scaling="weak",
                        config={"workload": "configs[4]: object_aware_icp, 8192 pts, n_slot 10",
                                "sample": "4 clouds x 2 ICP iterations on the host cores, scaled to icp_iter 20 (cost is linear in iterations)"},
                        cpu_baseline={"value": value * 2 / 20, "unit": "clouds/s", "cores": cores, "kind": "port",
                                      "sample": "4 clouds x 2 iterations, scaled to 20 iterations"})
        else:
            npoint = 4096 if cfg == "ogcdr_flow4096" else 2048
            value, ms = run_cpu_flow(npoint, 2, 4, max(1, min(args.steps, 2)), cores)
            line = dict(base, metric=f"pairs/sec ({npoint} pts, FlowStep3D iters 4 fwd + flow loss bwd + Adam)", value=value,
                        unit="pairs/s", ms_per_step=ms, scaling="weak",
                        config={"workload": f"configs[2]: flownet_ogcdr npoint {npoint} iters 4", "sample": "batch 2 (of 16) per step"},
                        cpu_baseline={"value": value, "unit": "pairs/s", "cores": cores, "kind": "port", "sample": "batch 2 per step"})
        line["e2e"] = {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
        emit(line)
        return

    if cfg == "sapien_cpu":       # configs[0] is CPU plumbing by definition: it only exists as the reference arm
        if rank == 0:
            emit(dict(cpu_arm("sapien_cpu", args.steps, args.warmup), note="configs[0] is a CPU-only plumbing case"))
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
    peaks = measured_peaks()
    peak_src = "MEASURED_PEAKS.json (measured)" if peaks["measured"] else "fallback 6650 GB/s / 1590 TFLOP/s"

    if cfg in ("oa_icp", "ogcdr_flow", "ogcdr_flow4096"):
        if rank != 0:                     # replicas only (no exchange step): one rank measures
            if world > 1:
                dist.destroy_process_group()
            return
        sampler = ClockSampler(local_rank)
        sampler.start()
        if cfg == "oa_icp":
            from ogc_b200 import icp
            B, IT = 64, 20
            host = [t.pin_memory() for t in icp_inputs(B, torch.device("cpu"))]
            dev_in = [t.to(device) for t in host]
            l0 = be.launches
            ms, _ = time_cuda(lambda i: icp.object_aware_icp(*dev_in, icp_iter=IT), args.steps, max(args.warmup, 3))
            launches = (be.launches - l0) / (args.steps + max(args.warmup, 3))
            out_host = torch.empty(B, N_POINT, 3).pin_memory()

            def e2e_fn(i):
                x = [t.to(device, non_blocking=True) for t in host]
                out_host.copy_(icp.object_aware_icp(*x, icp_iter=IT), non_blocking=True)
                torch.cuda.current_stream().synchronize()
            ms_e2e, _ = time_cuda(e2e_fn, args.steps, 1)
            pair_evals = B * N_POINT * N_POINT * IT
            line = {"metric": "point-clouds/sec (8192 pts, object-aware ICP, icp_iter 20)", "value": B / (ms * 1e-3),
                    "unit": "clouds/s", "ms_per_step": ms,
                    "config": {"workload": f"configs[4]: oa_icp.object_aware_icp, {B} clouds x {N_POINT} pts, n_slot {N_SLOT}, icp_iter {IT}",
                               "parallelism": "replicas only", "l2": "inputs 25 MB < L2; the kernel is ALU-bound (N^2 pair evaluations), not HBM-bound"},
                    "e2e": {"value": B / (ms_e2e * 1e-3), "unit": "clouds/s", "ms_per_step": ms_e2e,
                            "h2d_bytes_per_step": sum(t.numel() * 4 for t in host), "d2h_bytes_per_step": out_host.numel() * 4},
                    "roofline": {"kernel": "icp_correspond", "bound": "alu", "achieved": pair_evals / (ms * 1e-3) / 1e12,
                                 "unit": "T pair-evaluations/s", "peak": None, "frac": None, "traffic": None,
                                 "note": "streaming softmax over pc2 tiles in shared memory: ~14 fp32 ops + 1 ex2 per pair; "
                                         "algorithmic HBM bytes B*(N(36+4K)+N(12+4K)) = 67 MB per call are negligible"}}
            if not args.no_ref_ext:
                ref = ref_arm("oa_icp", "--clouds", B, "--icp-iter", IT, "--chunk", 4, "--steps", 2, "--warmup", 1, timeout=900)
                if "value" in ref:
                    ref["speedup_e2e"] = line["e2e"]["value"] / ref["value"]
                line["ref_cuda_ext"] = ref
        else:
            npoint = 4096 if cfg == "ogcdr_flow4096" else 2048
            batch, iters = 16, 4
            step, h2d = flow_step_fn(npoint, batch, iters, device)
            l0 = be.launches
            ms, last = time_cuda(lambda i: step(i, host_inputs=False), args.steps, max(args.warmup, 3))
            launches = (be.launches - l0) / (args.steps + max(args.warmup, 3))
            ms_e2e, last = time_cuda(step, args.steps, 1)        # pinned host inputs copied every step + loss read-back
            line = {"metric": f"pairs/sec ({npoint} pts, FlowStep3D iters 4 fwd + flow loss bwd + Adam)", "value": batch / (ms * 1e-3),
                    "unit": "pairs/s", "ms_per_step": ms,
                    "config": {"workload": f"configs[2]: flownet_ogcdr.FlowStep3D npoint {npoint}, batch {batch}, iters {iters} + "
                                           "UnsupervisedFlowStep3DLoss + backward + Adam (train_flow.py:59-92)",
                               "parallelism": "replicas only (BatchNorm statistics are replica-local)",
                               "l2": "launch-bound: ~25 FPS + ~30 kNN calls per forward on <= 4096 points"},
                    "e2e": {"value": batch / (ms_e2e * 1e-3), "unit": "pairs/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": h2d,
                            "d2h_bytes_per_step": 4 * len(last)},
                    "roofline": None, "loss": last}
            # per-kernel timing of OUR kernels over two eager steps (CUDA events on the launching stream)
            backend.TIMER.enabled = True
            backend.TIMER.reset()
            for i in range(2):
                step.trainer.train_step(step.resident[i % 2])
            torch.cuda.synchronize()
            backend.TIMER.enabled = False
            op_rows = summarise_ops(backend.TIMER.summary(), 2, peaks)
            line["ops"] = op_rows
            fam = [k for k in op_rows if k.startswith("flow_mlp")]
            if fam:
                top = max(fam, key=lambda k: op_rows[k]["ms_per_step"])
                r = op_rows[top]
                fp32_peak = 2 * 128 * 148 * 1965.0e6 / 1e12      # FFMA: 128 lanes x 2 flops per SM per clock at the 1965 MHz boost clock
                tr = {}
                tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
                if os.path.exists(tpath):
                    tr = json.load(open(tpath)).get(top, {})
                line["roofline"] = {"kernel": top, "bound": "hbm", "achieved": r["alg_gbs"], "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                    "frac": r["alg_gbs"] / peaks["hbm_gbs"], "peak_how": peak_src,
                                    "traffic": tr.get("dram_bytes_per_launch"),
                                    "traffic_note": ("ncu DRAM bytes per launch of this family in ONE representative block (16 x 67>64>64>64 x "
                                                     f"8192 positions), whose algorithmic bytes per launch are {tr['alg_bytes_same_launches']:.3g}: "
                                                     + tr.get("note", "")) if tr else None,
                                    "alg_bytes_per_launch": r["alg_bytes_per_launch"], "share_of_step": r["ms_per_step"] / ms,
                                    "fp32_simt_tflops": r.get("useful_tflops"), "fp32_simt_peak_tflops": fp32_peak,
                                    "fp32_simt_frac": (r.get("useful_tflops") or 0.0) / fp32_peak,
                                    "note": "the dominant family of the FlowStep3D blocks' shared MLPs: fp32 FFMA2 pointwise contractions at "
                                            "16-128 FLOP per algorithmic byte, a few hundred small launches per step (M*S = 2k-16k positions "
                                            "per sample): neither roofline is approached, the launches are latency / tail bound; both "
                                            "fractions reported"}
            # the same step with the dense inner layers on the tensor cores (3xTF32 split, csrc/sa_*_tma.cu): fp32-grade per
            # block (5e-7 of fp64) but outside the 1e-4 golden bound after two recurrent iterations -> labelled, not the headline
            del step
            torch.cuda.empty_cache()
            from ogc_b200 import bn_fused
            prev_tma, bn_fused.USE_TMA = bn_fused.USE_TMA, True
            try:
                step_tc, _ = flow_step_fn(npoint, batch, iters, device)
                ms_tc, _ = time_cuda(step_tc, args.steps, max(args.warmup, 3))
                line["tensor_core_inner_layers_3xtf32"] = {
                    "value": batch / (ms_tc * 1e-3), "unit": "pairs/s", "ms_per_step": ms_tc,
                    "note": "bn_fused.USE_TMA = True (OGC_BN_TMA=1): not the parity configuration (flow golden 1.3e-4 vs the 1e-4 bound)"}
                del step_tc
            finally:
                bn_fused.USE_TMA = prev_tma
            if not args.no_ref_ext:
                ref = ref_arm("flow", "--npoint", npoint, "--batch", batch, "--iters", iters, "--steps", 10, "--warmup", 3)
                if "value" in ref:
                    ref["speedup_e2e"] = line["e2e"]["value"] / ref["value"]
                line["ref_cuda_ext"] = ref
        line.update({"n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "higher_is_better": True,
                     "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                     "gpu_launches": launches, "clocks": sampler.stop()})
        if not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_arm(cfg, 2, 1).get("cpu_baseline", {"unavailable": "cpu arm failed"})
        emit(line)
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- segnet training step (kittisf / train8 / strong32) ----
    trainer = build_trainer(device, world, pairs)
    n_batches = 4
    batches = [data.make_batch(1000 * rank + i, pairs, N_POINT, aug=aug, fps_fn=be.fps, device=device)
               for i in range(n_batches)]
    resident = [tuple(x.to(device) for x in b) for b in batches]
    h2d = sum(x.numel() * x.element_size() for x in (batches[0][0], batches[0][2]))
    clouds_per_step = pairs * t_views * world
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
            if args.eager or args.no_prefetch:
                last = step_fn(it0 + i, src[i % n_batches], aug_transform=aug)
            else:       # the batch of the next call is known (a loader one batch ahead): its first-level FPS runs under this step
                last = step_fn(it0 + i, src[i % n_batches], aug_transform=aug, next_batch=src[(i + 1) % n_batches])
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
        if args.eager or args.no_prefetch:
            step_fn(it0 + i, resident[i % n_batches], aug_transform=aug)
        else:
            step_fn(it0 + i, resident[i % n_batches], aug_transform=aug, next_batch=resident[(i + 1) % n_batches])
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

    hbm_peak, tf32_peak = peaks["hbm_gbs"], peaks["bf16_tflops"] / 2.0
    op_rows = summarise_ops(ops, 2, peaks)
    # the dominant kernel family of the step's CRITICAL PATH: the latency-bound single-wave kernels that run on side
    # streams underneath it (FPS chain, three_nn, device Hungarian, nuclear norm: DESIGN.md "Scheduling") are listed in
    # `ops` with their own figures but do not set the roofline line
    hidden = {"fps", "three_nn", "mask_match", "mask_nuclear_norm", "mask_contingency"}
    cand = [k for k in op_rows if k not in hidden] or list(op_rows)
    top = max(cand, key=lambda k: op_rows[k]["ms_per_step"]) if cand else None
    roofline = None
    if top:
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get(top, {}).get("dram_bytes_per_launch")
        r = op_rows[top]
        alg = ops[top]["bytes"] / ops[top]["calls"]
        if r.get("bound") == "tensor":
            roofline = {"kernel": top, "bound": "tensor", "achieved": r["useful_tflops"], "peak": tf32_peak, "unit": "TFLOP/s",
                        "frac": r["useful_tflops"] / tf32_peak, "tensor_work_frac_3xtf32": 3 * r["useful_tflops"] / tf32_peak,
                        "peak_how": "TF32 dense = half the measured bf16 cuBLAS burst peak; " + peak_src,
                        "note": "useful fp32-equivalent FLOPs (2*Cin*Cout per position) per launch / CUDA-event duration; the "
                                "3xTF32 split issues three times that tensor work"}
        else:
            roofline = {"kernel": top, "bound": "hbm", "achieved": r["alg_gbs"], "peak": hbm_peak, "unit": "GB/s",
                        "frac": r["alg_gbs"] / hbm_peak, "peak_how": peak_src,
                        "note": "algorithmic bytes per launch / CUDA-event duration (DESIGN.md byte formulas)"}
            if "frac_of_tf32_peak" in r:
                roofline.update({"flop_per_byte_3xtf32": r["flop_per_byte_3xtf32"], "ridge_flop_per_byte": r["ridge_flop_per_byte"],
                                 "tensor_frac_useful": r["frac_of_tf32_peak"], "tensor_work_frac_3xtf32": r["frac_of_tf32_peak_3xtf32_work"],
                                 "note": "a contraction kernel below the ridge of the two measured peaks (3xTF32 work per algorithmic "
                                         "byte): HBM is its roofline; the tensor fractions are reported alongside"})
        roofline.update({"traffic": traffic, "alg_bytes_per_launch": alg,
                         "waste_ratio": (traffic / alg) if traffic and alg else None,
                         "share_of_step": r["ms_per_step"] / ms_step})
        value = clouds_per_step / (ms_step * 1e-3)
        roofline["step"] = {"alg_bytes_per_cloud": ALG_BYTES_PER_CLOUD, "bound": "hbm",
                            "frac": value / world * ALG_BYTES_PER_CLOUD / (hbm_peak * 1e9),
                            "note": "whole step against SURVEY 8d's perfectly-fused 42.5 MB per cloud"}

    line = {"metric": METRIC, "value": clouds_per_step / (ms_step * 1e-3), "unit": "clouds/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong" if cfg == "strong32" else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": seg_workload, "name": cfg, "parallelism": f"dp{world}",
                       "launch": "eager" if args.eager else "one CUDA graph per step" + (
                           "" if args.no_prefetch else "; the NEXT batch's clouds are copied in and their first-level FPS "
                           "centres sampled on a side stream of this step's graph (one copy and one FPS per step, as without it)"),
                       "l2": "per-step working set (GBs of activations) >> 126 MB L2; inputs cycle over 4 distinct batches"},
            "e2e": {"value": clouds_per_step / (ms_e2e * 1e-3), "unit": "clouds/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4 * len(last_dict)},
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "ops": op_rows,
            "loss": last_dict}

    if world > 1:
        dist.destroy_process_group()
    # free this process's GPU memory before the subprocess legs (the reference arm needs ~19 GB)
    del trainer, resident, batches
    torch.cuda.empty_cache()
    if world == 1 and not args.no_ref_ext:
        refs = {}
        for mode, flag in (("tf32_default", []), ("strict_fp32", ["--fp32"])):
            refs[mode] = ref_arm("kittisf", "--steps", 10, "--warmup", 3, "--pairs", pairs, *(["--no-aug"] if not aug else []), *flag)
        main_ref = refs["tf32_default"]
        ref = dict(main_ref)
        if "value" in ref:
            ref["speedup_e2e"] = line["e2e"]["value"] / ref["value"]
        if "value" in refs["strict_fp32"]:
            ref["strict_fp32"] = {k: refs["strict_fp32"][k] for k in ("value", "ms_per_step", "steps", "precision")}
            ref["speedup_e2e_vs_strict_fp32"] = line["e2e"]["value"] / refs["strict_fp32"]["value"]
        line["ref_cuda_ext"] = ref
    if world == 1 and not args.no_cpu_baseline:
        cpu = cpu_arm(cfg, 2, 1)                    # 2 steps (+1 warm-up) of the SAME 16-cloud step: ~25 s of CPU work
        line["cpu_baseline"] = cpu.get("cpu_baseline", cpu)
    emit(line)


if __name__ == "__main__":
    main()
