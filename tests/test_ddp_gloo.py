"""Data-parallel step on CPU (gloo, world_size 2): the sharded step with ONE all-reduce over
[flat grads | NaN counter] must equal the single-process step on the concatenated batch
(clouds are independent, loss = mean over samples)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _build(world, seed=10):
    from ogc_b200 import losses
    from ogc_b200.segnet import MaskFormer3D
    from ogc_b200.train import SegTrainer
    torch.manual_seed(seed)
    net = MaskFormer3D(n_slot=6, n_point=256, variant="sapien")
    cfg = {**losses.KITTISF_LOSS_CFG, "start_steps": [0, 0, 0],
           "smooth_loss_params": {"w_knn": 3.0, "w_ball_q": 1.0,
                                  "knn_loss_params": {"k": 8, "radius": 0.1, "loss_norm": 1},
                                  "ball_q_loss_params": {"k": 16, "radius": 0.2, "loss_norm": 1}}}
    return SegTrainer(net, losses.build_ogc_loss(cfg), global_batch_size=2, world_size=world)


def _batch():
    from ogc_b200 import data
    pcs, segms, flows, valids = data.make_batch(3, 2, 256, aug=True)
    return pcs * 0.02, segms, flows * 0.02, valids


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    from ogc_b200 import backend
    from oracle.pointnet2_oracle import OracleBackend
    backend.set_backend(OracleBackend())
    # rank 1 initialises from a DIFFERENT seed: SegTrainer must broadcast rank 0's weights at start-up (ADVICE r1:
    # replicas that only ever all-reduce gradients silently train different models otherwise)
    tr = _build(world, seed=10 + 7 * rank)
    full = _batch()
    shard = tuple(x[rank:rank + 1] for x in full)
    d = tr.train_step(5000, shard, aug_transform=True)
    d2 = tr.train_step(5001, shard, aug_transform=True)
    if rank == 0:
        torch.save({"p": tr.opt.flat_p.clone(), "g": tr.opt.flat_g.clone(), "nan": tr.opt.nan_counter.clone(),
                    "d": d, "d2": d2}, out)
    dist.destroy_process_group()


def test_two_rank_step_equals_single_process(tmp_path, oracle):
    port, out = _free_port(), str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = torch.load(out)

    from ogc_b200 import backend
    prev = backend.set_backend(oracle)
    try:
        torch.set_num_threads(4)
        tr = _build(1)
        full = _batch()
        d = tr.train_step(5000, full, aug_transform=True)
        tr.train_step(5001, full, aug_transform=True)
    finally:
        backend.set_backend(prev)
    # after the all-reduce the buffer holds the SUM of the two ranks' gradients = 2 x the global-mean gradient
    torch.testing.assert_close(got["g"] / 2, tr.opt.flat_g, rtol=2e-3, atol=2e-6)
    torch.testing.assert_close(got["p"], tr.opt.flat_p, rtol=0, atol=2e-4)
    assert float(got["nan"]) == 0.0
    # rank 0's logged loss is its shard's; the global loss is the mean of the two shards (checked via grads)
    assert abs(got["d"]["sum"] - d["sum"]) < 0.5 * abs(d["sum"]) + 1.0


def test_nan_gradient_skips_update_on_every_rank():
    """train_seg.py:81-83: a NaN anywhere -> no optimizer step (here agreed through the all-reduced counter)."""
    from ogc_b200.train import FlatAdam
    p = torch.nn.Parameter(torch.ones(5))
    opt = FlatAdam([p], lr=0.1)
    opt.zero_grad()
    p.grad.copy_(torch.tensor([1.0, float("nan"), 0, 0, 0]))
    before = opt.flat_p.clone()
    opt.step()
    assert torch.equal(opt.flat_p, before)
    assert opt.t == 0          # a skipped step does not advance the bias correction (the reference never calls step())
    opt.zero_grad()
    p.grad.fill_(1.0)
    opt.step()
    assert (opt.flat_p < before).all()
