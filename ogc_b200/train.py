"""The benchmarked training step -- host-side mirror of `Trainer._train_it` (train_seg.py:47-86) for the
data-parallel hot path, one process per GPU.

    batch (b,t,N,3) --H2D--> segnet fwd --> UnsupervisedOGCLoss --> backward --> NaN guard --> Adam

Differences from the reference (all behaviour-preserving):
  * parameters, gradients and the Adam moments live in three flat fp32 buffers; `zero_grad` is one
    memset, the NaN scan (train_seg.py:81-83: one host sync per parameter) is one kernel leaving a
    device-side counter, `optimizer.step()` is one kernel that is skipped ON THE DEVICE when the
    counter is non-zero (csrc/optim.cu) -- no host sync between backward and the update;
  * multi-GPU (the reference has none): clouds are independent, so ranks hold disjoint
    sample shards and exchange exactly ONE NCCL all-reduce per step over [flat grads | NaN counter];
    the 1/world_size scale is folded into the Adam kernel.  The step-weight schedule and `lr_curve`
    use the GLOBAL batch size (train_seg.py:70, :230-234).
"""
import ctypes

import torch
import torch.distributed as dist

from . import _lib
from .backend import TIMER, get_backend


def lr_curve(it, batch_size, lr, lr_decay, decay_step, lr_clip):
    """train_seg.py:230-234"""
    return max(lr_decay ** (int(it * batch_size / decay_step)), lr_clip / lr)


class FlatAdam:
    """Adam over one flat buffer; parameters and their .grad are re-pointed to views of it."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        self.params = [p for p in params if p.requires_grad]
        dev = self.params[0].device
        n = sum(p.numel() for p in self.params)
        self.n = n
        self.flat_p = torch.empty(n, dtype=torch.float32, device=dev)
        # gradient buffer carries one extra float: the NaN counter rides the same all-reduce
        self.flat_g_ext = torch.zeros(n + 1, dtype=torch.float32, device=dev)
        self.flat_g = self.flat_g_ext[:n]
        self.nan_counter = self.flat_g_ext[n:]
        self.m = torch.zeros(n, dtype=torch.float32, device=dev)
        self.v = torch.zeros(n, dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:
            k = p.numel()
            self.flat_p[off:off + k].copy_(p.data.reshape(-1))
            p.data = self.flat_p[off:off + k].view_as(p.data)
            p.grad = self.flat_g[off:off + k].view_as(p.data)
            off += k
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.t = 0
        self.lib = _lib.load() if dev.type == "cuda" else None
        if self.lib is not None:
            # step counter + learning rate live on the device so a captured step can be replayed
            self.state = torch.zeros(2, dtype=torch.float32, device=dev)
            self.lr_host = torch.zeros(1, dtype=torch.float32).pin_memory()

    def zero_grad(self):
        self.flat_g_ext.zero_()

    def step(self, lr_scale=1.0, grad_scale=1.0):
        if self.lib is None:      # CPU arm of bench.py (--impl reference): same arithmetic in torch
            if bool(torch.isnan(self.flat_g).any()):
                return            # skipped step: the bias-correction counter does not advance (train_seg.py:81-85)
            self.t += 1
            g = self.flat_g * grad_scale
            if self.weight_decay:
                g = g + self.weight_decay * self.flat_p
            b1, b2 = self.betas
            self.m.mul_(b1).add_(g, alpha=1 - b1)
            self.v.mul_(b2).addcmul_(g, g, value=1 - b2)
            denom = self.v.sqrt() / (1 - b2 ** self.t) ** 0.5 + self.eps
            self.flat_p.addcdiv_(self.m, denom, value=-self.lr * lr_scale / (1 - b1 ** self.t))
            return
        self.t += 1
        self.set_lr(self.lr * lr_scale)
        self.launch_step(grad_scale)

    def set_lr(self, lr):
        """Host -> device copy of this step's learning rate (outside any CUDA graph)."""
        self.lr_host[0] = lr
        self.state[1:2].copy_(self.lr_host, non_blocking=True)

    def launch_step(self, grad_scale=1.0):
        """The update kernels only (capturable): t += 1 on the device, then Adam with state[1] as lr."""
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        P = lambda t: ctypes.c_void_p(t.data_ptr())
        with TIMER.span("adam_step", 28 * self.n):
            _lib.check(self.lib.ogc_adam_step_dev(self.n, P(self.flat_p), P(self.flat_g), P(self.m), P(self.v),
                                                  P(self.state), self.betas[0], self.betas[1], self.eps,
                                                  self.weight_decay, grad_scale, P(self.nan_counter), st),
                       "ogc_adam_step_dev")
        get_backend().launches += 2

    def count_nan(self):
        if self.lib is None:
            return
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        with TIMER.span("count_nan", 4 * self.n):
            _lib.check(self.lib.ogc_count_nan(self.n, ctypes.c_void_p(self.flat_g.data_ptr()),
                                              ctypes.c_void_p(self.nan_counter.data_ptr()), st), "ogc_count_nan")
        get_backend().launches += 1


class SegTrainer:
    def __init__(self, segnet, criterion, lr=1e-3, weight_decay=0.0, lr_decay=0.7, lr_clip=1e-5,
                 decay_step=200000, global_batch_size=4, world_size=1):
        self.segnet, self.criterion = segnet, criterion
        self.opt = FlatAdam(segnet.parameters(), lr=lr, weight_decay=weight_decay)
        self.sched = dict(lr=lr, lr_decay=lr_decay, decay_step=decay_step, lr_clip=lr_clip)
        self.global_batch_size, self.world_size = global_batch_size, world_size
        self.device = self.opt.flat_p.device
        if world_size > 1:
            self.sync_replicas()
        self.overlap_geometry = True      # FPS chain on a side stream under the loss neighbourhoods
        self._geo_stream = None
        self._geo_stream2 = None

    def sync_replicas(self, src=0):
        """Data-parallel replicas must start from (and resume with) identical weights and optimiser state: broadcast
        rank `src`'s flat parameter / moment buffers and step counter.  Called at construction; call it again after
        loading a checkpoint on one rank.  (The per-step all-reduce only keeps EQUAL replicas equal.)"""
        if self.world_size <= 1 or not dist.is_initialized():
            return
        for buf in (self.opt.flat_p, self.opt.m, self.opt.v):
            dist.broadcast(buf, src=src)
        if self.opt.lib is not None:
            dist.broadcast(self.opt.state, src=src)
        t = torch.tensor([float(self.opt.t)], device=self.device)
        dist.broadcast(t, src=src)
        self.opt.t = int(t.item())
        # cheap divergence check: every rank must now hold the same parameter checksum
        chk = self.opt.flat_p.double().sum().reshape(1)
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        if float(hi - lo) != 0.0:
            raise RuntimeError("ogc_b200: replicas disagree on the parameters after the start-up broadcast")

    def _step_body(self, pcs, flows, it, aug_transform, defer, allreduce=True, c1=None):
        """zero_grad -> forward -> loss -> backward -> NaN count -> all-reduce -> Adam launch (no host sync when
        `defer`).  pcs, flows: (b,t,N,3) on the device."""
        self.segnet.train()
        self.opt.zero_grad()
        b, t, n, _ = pcs.shape
        flat = pcs.view(b * t, n, 3)
        pcs_l = [pcs[:, i].contiguous() for i in range(t)]
        centres, fp_nn = (self._prefetch_geometry(flat, pcs_l, c1) if self.overlap_geometry and flat.is_cuda
                          else (None, None))
        masks = self.segnet(flat, flat, centres, fp_nn).view(b, t, n, -1)
        if centres is not None and getattr(self, "_geo_join", None):
            # every side-stream product has been consumed through its event; join the streams themselves as well so
            # that a stream capture sees no unjoined branch
            for st in self._geo_join:
                if st is not None:
                    torch.cuda.current_stream().wait_stream(st)
            self._geo_join = None
        masks_l = [masks[:, i].contiguous() for i in range(t)]
        flows_l = [flows[:, i].contiguous() for i in range(t)]
        self.criterion.defer_logging = defer
        try:
            loss, loss_dict = self.criterion(pcs_l, masks_l, flows_l, step_w=True, it=it * self.global_batch_size,
                                             aug_transform=aug_transform)
        finally:
            self.criterion.defer_logging = False
        from . import sa_fused
        sa_fused.ACCUMULATE_INTO_GRAD = True       # the fused blocks' kernels accumulate into the flat gradient buffer directly
        try:
            loss.backward()
        finally:
            sa_fused.ACCUMULATE_INTO_GRAD = False
        self.opt.count_nan()
        if self.world_size > 1 and allreduce:
            dist.all_reduce(self.opt.flat_g_ext)          # the step's only collective: grads + NaN counter
        return loss_dict

    def _prefetch_geometry(self, flat, pcs_l, c1=None):
        """Everything that depends on the coordinates alone, arranged so that the latency-bound FPS chain (one CTA
        per cloud: 16 of 148 SMs busy for ~1.8 ms at KITTI-SF sizes) runs on a side stream UNDERNEATH the loss
        neighbourhoods (k-NN + ball query on 8192 x 8192, thousands of small CTAs that flow around it).  Fork / join
        with stream events, so it is captured into the step's CUDA graph as two parallel branches."""
        from . import losses
        be = get_backend()
        if getattr(be, "name", "") != "b200" or losses.FORCE_COMPOSED:
            return None, None
        main = torch.cuda.current_stream()
        if self._geo_stream is None:
            self._geo_stream = torch.cuda.Stream()
        side = self._geo_stream
        side.wait_stream(main)
        from . import segnet as _segnet
        third = None
        with torch.cuda.stream(side):
            if _segnet.FORCE_COMPOSED:
                centres, fp_nn = self.segnet.sample_chain(flat), None
            else:
                # the three_nn of the finest FP level (8192 <- 2048, the only sizeable one) needs the first level's
                # centres only: it runs on a third stream next to the rest of the (16-CTA) FPS chain
                if self._geo_stream2 is None:
                    self._geo_stream2 = torch.cuda.Stream()
                third = self._geo_stream2
                sa = self.segnet.SA_modules
                # every product carries the event it becomes ready at (`_ogc_ready`): the main stream waits per
                # CONSUMER (level-1 centres before SA1, ... -- segnet.wait_ready), not for the whole chain, so that
                # only the first FPS (1.27 ms of the 1.77 ms chain) sits on the critical path
                def ready(t, stream):
                    t._ogc_ready = torch.cuda.Event()
                    t._ogc_ready.record(stream)
                    return t
                # c1: the first level's centres when they were sampled during the PREVIOUS step (train_step_graphed with
                # next_batch): the 1.3 ms single-wave FPS of the input clouds leaves the critical path altogether
                centres = [ready(sa[0].sample(flat), side) if c1 is None else ready(c1, side)]
                with torch.cuda.stream(third):
                    third.wait_event(centres[0]._ogc_ready)
                    nn0 = be.three_nn(flat.contiguous(), centres[0])
                    ready(nn0[0], third)
                for m in sa[1:]:
                    centres.append(ready(m.sample(centres[-1]), side))
                l_pc = [flat] + centres
                fp_nn = [nn0]
                for i in range(1, len(self.segnet.FP_modules)):
                    nn_i = be.three_nn(l_pc[i].contiguous(), l_pc[i + 1].contiguous())
                    ready(nn_i[0], side)
                    fp_nn.append(nn_i)
        specs = losses.smooth_specs(self.criterion.smooth_loss) if hasattr(self.criterion, "smooth_loss") else None
        losses.NEIGHBOUR_CACHE.clear()
        nbs = None
        if specs:
            # the loss neighbourhoods are consumed after the network's forward: with the first-level centres prefetched
            # (c1) nothing hides them any more, so they get a stream of their own and are joined before the criterion
            if c1 is not None and third is not None:
                if getattr(self, "_nb_stream", None) is None:
                    self._nb_stream = torch.cuda.Stream()
                nbs = self._nb_stream
                nbs.wait_stream(main)
            with torch.cuda.stream(nbs if nbs is not None else main):
                for pc in pcs_l:
                    handle = losses.tag_cloud(pc)
                    for kind, k, radius in specs:
                        losses.NEIGHBOUR_CACHE[(handle, kind, k, radius)] = losses.neighbourhood(be, kind, k, radius, pc)
        if third is None:                 # composed path: plain join
            main.wait_stream(side)
        self._geo_join = (side, third, nbs)
        return centres, fp_nn

    def train_step(self, it, batch, aug_transform=False):
        """batch = (pcs (b,t,N,3), segms, flows (b,t,N,3), valids) on host or device.  Returns loss_dict.
        Eager path (one launch per kernel); see train_step_graphed for the CUDA-graph replay."""
        lr_scale = lr_curve(it, self.global_batch_size, **self.sched)
        pcs, _, flows, _ = batch
        pcs = pcs.to(self.device, non_blocking=True)
        flows = flows.to(self.device, non_blocking=True)
        loss_dict = self._step_body(pcs, flows, it, aug_transform, defer=False)
        self.opt.step(lr_scale=lr_scale, grad_scale=1.0 / self.world_size)
        return loss_dict

    # ------------------------------------------------------------------------------------------------
    # CUDA-graph replay: the whole step (H2D copies excluded) is ONE graph launch.  Possible because the
    # Hungarian matching, the NaN guard, the Adam step counter and the learning rate all live on the device.
    # ------------------------------------------------------------------------------------------------
    def _loss_weights(self, it):
        c = self.criterion
        k = it * self.global_batch_size
        return (c.step_lossw(k, c.w_dynamic, c.start_step_dynamic), c.step_lossw(k, c.w_smooth, c.start_step_smooth),
                c.step_lossw(k, c.w_invariance, c.start_step_invariance))

    def _capture(self, key, it, shape, aug_transform):
        from .losses import resolve_loss_dict  # noqa: F401
        dev = self.device
        g = {"pcs": torch.zeros(shape, dtype=torch.float32, device=dev),
             "flows": torch.zeros(shape, dtype=torch.float32, device=dev)}
        g["pcs"].copy_(self._last_inputs[0]); g["flows"].copy_(self._last_inputs[1])
        prefetch = key[-1]
        if prefetch:
            # software pipeline across steps: this step's graph samples the first-level centres of the NEXT step's clouds
            # on a side stream (a function of the coordinates alone: same kernel, same input, same result as sampling
            # them inside the next step), and starts from the centres the previous step left in c1_next
            sa0 = self.segnet.SA_modules[0]
            g["next_pcs"] = g["pcs"].clone()
            g["c1_cur"] = sa0.sample(g["pcs"].view(-1, shape[2], 3))
            g["c1_next"] = g["c1_cur"].clone()
            g["fps_stream"] = torch.cuda.Stream()

        def body(allreduce):
            if not prefetch:
                return self._step_body(g["pcs"], g["flows"], it, aug_transform, defer=True, allreduce=allreduce)
            main = torch.cuda.current_stream()
            g["c1_cur"].copy_(g["c1_next"])
            g["fps_stream"].wait_stream(main)
            with torch.cuda.stream(g["fps_stream"]):
                g["c1_next"].copy_(self.segnet.SA_modules[0].sample(g["next_pcs"].view(-1, shape[2], 3)))
            d = self._step_body(g["pcs"], g["flows"], it, aug_transform, defer=True, allreduce=allreduce, c1=g["c1_cur"])
            main.wait_stream(g["fps_stream"])
            return d
        opt = self.opt
        snap = [x.clone() for x in (opt.flat_p, opt.m, opt.v, opt.state)]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                      # warm-up on a side stream (allocator, cuBLAS, NCCL)
            for _ in range(2):
                # no collective in the warm-up: a rank-local re-capture (e.g. a ragged last batch on one rank) must not
                # change the number of all-reduces the ranks issue
                body(False)
                opt.launch_step(1.0 / self.world_size)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        for dst, src in zip((opt.flat_p, opt.m, opt.v, opt.state), snap):   # warm-up must not train
            dst.copy_(src)
        launches0 = get_backend().launches
        host = torch.zeros(16, dtype=torch.float32).pin_memory()      # allocated BEFORE capture
        # world_size > 1: the graph ends before the collective; the NCCL all-reduce, the Adam kernel and the D2H of the
        # logged scalars follow it eagerly on the same stream (3 launches) -- no NCCL work inside a stream capture.
        multi = self.world_size > 1
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            d = body(False)
            g["keys"] = d["_keys"]
            g["values"] = d["_values"]
            g["host"] = host[:len(d["_keys"])]
            if not multi:
                opt.launch_step(1.0)
                g["host"].copy_(d["_values"], non_blocking=True)       # captured D2H of the logged scalars
        g["launches"] = get_backend().launches - launches0
        g["graph"] = graph
        # keep ONLY the current key: the start_steps schedule and the augmentation switch are monotone, so an old
        # graph (multi-GB private activation pool + static inputs) is never replayed again
        self._graphs.clear()
        self._graphs[key] = g
        torch.cuda.empty_cache()
        return g

    def train_step_graphed(self, it, batch, aug_transform=False, next_batch=None):
        """Same semantics as train_step; the device work is a single CUDA-graph launch.  Re-captures when the
        batch shape or the active loss weights (start_steps schedule) change.
        next_batch: the batch of the FOLLOWING call, if the caller knows it (a data loader one batch ahead).  Its clouds
        are copied in now and their first-level FPS centres are sampled on a side stream of this step's graph; the
        following call then starts its network immediately.  The call after a `next_batch=None` call (or with another
        batch than announced) samples its centres eagerly first, so results never depend on the hint."""
        if not hasattr(self, "_graphs"):
            self._graphs = {}
        pcs, _, flows, _ = batch
        self._prefetch_mode = getattr(self, "_prefetch_mode", False) or next_batch is not None
        key = (tuple(pcs.shape), bool(aug_transform), self._loss_weights(it), self._prefetch_mode)
        self._last_inputs = (pcs, flows)
        fresh = key not in self._graphs
        g = self._graphs.get(key) or self._capture(key, it, tuple(pcs.shape), aug_transform)
        if self._prefetch_mode:
            if not fresh and g.get("announced") is pcs:
                g["pcs"].copy_(g["next_pcs"])              # already on the device, centres in c1_next
            else:
                g["pcs"].copy_(pcs, non_blocking=True)
                g["c1_next"].copy_(self.segnet.SA_modules[0].sample(g["pcs"].view(-1, pcs.shape[2], 3)))
            g["announced"] = None
            if next_batch is not None and tuple(next_batch[0].shape) == tuple(pcs.shape):
                g["next_pcs"].copy_(next_batch[0], non_blocking=True)
                g["announced"] = next_batch[0]
        else:
            g["pcs"].copy_(pcs, non_blocking=True)         # H2D from pinned memory (or D2D when resident)
        g["flows"].copy_(flows, non_blocking=True)
        self.opt.set_lr(self.opt.lr * lr_curve(it, self.global_batch_size, **self.sched))
        g["graph"].replay()
        get_backend().launches += g["launches"]
        if self.world_size > 1:
            dist.all_reduce(self.opt.flat_g_ext)           # the step's only collective: grads + NaN counter
            self.opt.launch_step(1.0 / self.world_size)
            g["host"].copy_(g["values"], non_blocking=True)
        torch.cuda.current_stream().synchronize()          # the logged scalars are in pinned host memory now
        out = dict(zip(g["keys"], g["host"].tolist()))
        out.setdefault("invariance", 0)
        return out


class FlowTrainer:
    """`Trainer._train_it` of train_flow.py:59-92 (FlowStep3D forward over `iters` GRU iterations, unsupervised flow loss,
    backward, NaN guard, Adam) with the same device-side machinery as SegTrainer: flat Adam with the NaN-skip decided on
    the device, and the whole step replayed as ONE CUDA graph (`train_step_graphed`) -- the reference's step is ~3000
    small launches and is bound by the host, not the GPU.  Replicas only across GPUs (BatchNorm statistics are local)."""

    def __init__(self, flownet, criterion, iters, lr=1e-3, weight_decay=0.0):
        self.net, self.criterion, self.iters = flownet, criterion, iters
        self.opt = FlatAdam(flownet.parameters(), lr=lr, weight_decay=weight_decay)
        self.device = self.opt.flat_p.device
        self._graph = None

    def _body(self, pcs, defer):
        self.net.train()
        self.opt.zero_grad()
        pc1, pc2 = pcs[:, 0].contiguous(), pcs[:, 1].contiguous()
        preds = self.net(pc1, pc2, pc1, pc2, iters=self.iters)
        self.criterion.defer_logging = defer
        try:
            loss, d = self.criterion(pc1, pc2, preds)
        finally:
            self.criterion.defer_logging = False
        loss.backward()
        self.opt.count_nan()
        return d

    def train_step(self, batch):
        pcs = batch[0].to(self.device, non_blocking=True)
        d = self._body(pcs, defer=False)
        self.opt.step()
        return d

    def train_step_graphed(self, batch):
        pcs = batch[0]
        if self._graph is None or self._graph["pcs"].shape != pcs.shape:
            g = {"pcs": torch.zeros(pcs.shape, dtype=torch.float32, device=self.device)}
            g["pcs"].copy_(pcs)
            opt = self.opt
            snap = [x.clone() for x in (opt.flat_p, opt.m, opt.v, opt.state)]
            bn = [(m, m.running_mean.clone(), m.running_var.clone(), m.num_batches_tracked.clone())
                  for m in self.net.modules() if isinstance(m, torch.nn.modules.batchnorm._BatchNorm)]
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):                      # warm-up (allocator, cuBLAS handles) on a side stream
                for _ in range(2):
                    self._body(g["pcs"], defer=True)
                    opt.launch_step(1.0)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            for dst, src in zip((opt.flat_p, opt.m, opt.v, opt.state), snap):        # the warm-up must not train
                dst.copy_(src)
            for m, a, b_, c in bn:
                m.running_mean.copy_(a); m.running_var.copy_(b_); m.num_batches_tracked.copy_(c)
            host = torch.zeros(32, dtype=torch.float32).pin_memory()
            l0 = get_backend().launches
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                d = self._body(g["pcs"], defer=True)
                opt.launch_step(1.0)
                g["keys"], g["host"] = d["_keys"], host[:len(d["_keys"])]
                g["host"].copy_(d["_values"], non_blocking=True)
            g["launches"] = get_backend().launches - l0
            g["graph"] = graph
            self._graph = g
        g = self._graph
        g["pcs"].copy_(pcs, non_blocking=True)
        self.opt.set_lr(self.opt.lr)
        g["graph"].replay()
        get_backend().launches += g["launches"]
        torch.cuda.current_stream().synchronize()
        return dict(zip(g["keys"], g["host"].tolist()))
