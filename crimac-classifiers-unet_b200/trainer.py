"""Data-parallel training loop body for the native U-Net (reference pipeline_train_predict/pipeline.py:144-203:
SGD(lr, momentum) + ExponentialLR, one optimizer step per batch).

One process per GPU.  Parameters and gradients live in two flat fp32 arenas.  The gradient exchange is the library's own
NVLink peer-memory all-reduce (csrc/peer_allreduce.cu), launched per gradient bucket inside backward and followed by the
fused SGD update of that bucket; `exchange="nccl"` keeps one ncclAllReduce of the whole arena after backward as the A/B
baseline.  The path has no other collective: BatchNorm statistics stay per replica, as in the reference (no SyncBN).  Semantics = DistributedDataParallel: the
update uses the mean over replicas of the per-replica (weighted-mean) loss gradients.  Like DDP, constructing a
Trainer inside an initialised process group broadcasts rank 0's parameters AND BatchNorm buffers to every replica.
Unlike DDP (broadcast_buffers=True) the BatchNorm running statistics are NOT re-synchronised on every forward: each
replica keeps the statistics of its own shard (the reference has no multi-GPU path to mirror); save rank 0's
state_dict, or call `broadcast_parameters()` before a checkpoint if all ranks must write identical files.
"""
import torch
import torch.distributed as dist

from . import engine as _engine


class PeerGradientExchange:
    """Symmetric-memory plumbing of the peer all-reduce (csrc/peer_allreduce.cu): allocates the flat gradient arena and
    the signal pad with torch.distributed._symmetric_memory (cuMemMap'ed on every GPU of the node, NVSwitch multicast
    address when the fabric offers one), exchanges the handles, and builds the crimac_comm_config every native train
    context of the model gets.  One process per GPU, NCCL process group already initialised."""

    def __init__(self, total_floats, device, group=None, use_multicast=None, ctas=0):
        import os
        import torch.distributed._symmetric_memory as symm_mem
        group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world > 8:
            raise RuntimeError("the peer-memory gradient exchange covers one NVSwitch node (<= 8 replicas)")
        padded = (total_floats + 1023) // 1024 * 1024
        self.arena = symm_mem.empty(padded, dtype=torch.float32, device=device)
        self.pad = symm_mem.empty(_engine.AR_PAD_BYTES // 4, dtype=torch.int32, device=device)
        self.arena.zero_()
        self.pad.zero_()
        h_arena = symm_mem.rendezvous(self.arena, group)
        h_pad = symm_mem.rendezvous(self.pad, group)
        self.state = torch.zeros(_engine.AR_STATE_BYTES // 4, dtype=torch.int32, device=device)
        if use_multicast is None:
            use_multicast = os.environ.get("CRIMAC_AR_MULTICAST", "1") != "0"
        mc = 0
        try:
            if use_multicast and h_arena.has_multicast_support(device.type, device.index if device.index is not None else 0):
                mc = int(h_arena.multicast_ptr)
        except Exception:
            mc = 0
        self.multicast = mc != 0
        self.comm = _engine.make_comm_config(self.world, self.rank, list(h_arena.buffer_ptrs), list(h_pad.buffer_ptrs),
                                             mc, self.state.data_ptr(), padded, ctas)
        self._handles = (h_arena, h_pad)
        torch.cuda.synchronize(device)
        dist.barrier(group)          # every pad is zero and mapped before the first flag is written


def gradient_buckets(model):
    """The three slices of the flat gradient arena (parameters() order) the native backward all-reduces as it goes, in the
    order they become final: (name, offset, count) in floats.  Bucket 0 = decoder blocks + head (the tail of the arena),
    1 = the two deepest encoder blocks, 2 = the remaining encoder blocks; the last one is padded to the arena's
    1024-float granularity.  Mirrors close_bucket() in csrc/net_api.cu (which derives the same ranges from the gradient
    pointers); used by the tests and for documentation."""
    sizes, off = {}, 0
    for name, p in model.named_parameters():
        sizes[name] = (off, p.numel())
        off += p.numel()
    total, depth = off, model.depth
    padded = (total + 1023) // 1024 * 1024
    split = depth - 2 if depth >= 3 else 0
    first_up = sizes["up_convs.0.upconv.weight" if model.up_mode == "transpose" else "up_convs.0.upconv.1.weight"][0]
    first_deep = sizes[f"down_convs.{split}.main.0.weight"][0]
    out = [("decoder+head", first_up, padded - first_up), ("deep encoder", first_deep, first_up - first_deep)]
    if split > 0:
        out.append(("shallow encoder", 0, first_deep))
    return out


def reduce_gradients(flat_grads, world):
    """Sum the flat gradient arena over all replicas (NCCL on GPUs, gloo in the CPU tests) and return the factor that
    turns the sum into the DDP mean; the factor is folded into the SGD kernel instead of a separate scaling pass."""
    if world > 1:
        dist.all_reduce(flat_grads, op=dist.ReduceOp.SUM)
    return 1.0 / world


class Trainer:
    """One optimisation step per call (`step`, `fit_host`, `fit_survey`).  Binding a Trainer to a model fuses the
    SGD(momentum) update into the model's native train step: `model.train_step_fused(...)` then ALSO updates the
    parameters (the autograd path `model(x)` / `loss.backward()` never does)."""

    def __init__(self, model, lr=0.005, momentum=0.95, lr_reduction=0.5, lr_step=1000, class_weight=(10.0, 300.0, 250.0),
                 use_cuda_graph=None, exchange=None, exchange_ctas=0, fused_optimizer=True):
        # defaults: reference configs/config_baseline.yaml:28-31,38 and pipeline.py:135
        self.model = model
        self.lr, self.momentum = float(lr), float(momentum)
        self.lr_reduction, self.lr_step = float(lr_reduction), int(lr_step)
        self.iteration = 0
        params = list(model.parameters())
        dev = params[0].device
        total = sum(p.numel() for p in params)
        self.flat_params = torch.empty(total, dtype=torch.float32, device=dev)
        off = 0
        for p in params:  # re-seat every parameter as a view of the arena (values preserved)
            n = p.numel()
            self.flat_params[off:off + n].copy_(p.data.reshape(-1))
            p.data = self.flat_params[off:off + n].view_as(p)
            off += n
        self.flat_momentum = torch.zeros_like(self.flat_params)
        self.class_weight = torch.tensor(class_weight, dtype=torch.float32, device=dev)
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        # The ~190 kernel launches of a step (both streams of the backward, the all-reduce, the SGD kernel) are captured
        # once in a CUDA graph and replayed: measured 3-4 % shorter steps on B200 (launch gaps).  The first step of a
        # shape runs eagerly (it creates the native context and the gradient arena), the second is captured; a change
        # of learning rate or batch shape re-captures.  use_cuda_graph=False keeps eager launches.
        if use_cuda_graph is None:
            use_cuda_graph = dev.type == "cuda"
        self.use_cuda_graph = bool(use_cuda_graph)
        self._graph = None
        self._graph_key = None
        self._eager_done = set()
        # gradient exchange: "peer" = our own NVLink peer-memory all-reduce kernels, launched bucket by bucket inside
        # backward and captured in the step's CUDA graph; "nccl" = one ncclAllReduce of the whole arena after backward
        # (kept as the A/B baseline: Trainer(..., exchange="nccl"))
        self.exchange = None
        if self.world > 1:
            self.exchange = exchange if exchange is not None else ("peer" if dev.type == "cuda" else "nccl")
            if self.exchange == "peer":
                # symmetric memory needs peer access between all GPUs of the process group (one NVSwitch node); if the
                # platform cannot provide it, every rank falls back - loudly, and together - to one ncclAllReduce per step
                try:
                    self.peer = PeerGradientExchange(total, dev, ctas=exchange_ctas)
                    ok = torch.ones(1, device=dev)
                except Exception as exc:          # noqa: BLE001 - any failure of the symmetric-memory plumbing
                    import warnings
                    warnings.warn(f"peer-memory gradient exchange unavailable ({exc!r}); using ncclAllReduce after backward")
                    self.peer = None
                    ok = torch.zeros(1, device=dev)
                dist.all_reduce(ok, op=dist.ReduceOp.MIN)
                if ok.item() < 1:
                    self.peer, self.exchange = None, "nccl"
                else:
                    model._grad_arena = self.peer.arena[:total]       # train_step_fused seats every p.grad in here
                    model._set_native_comm(self.peer.comm)
            self.broadcast_parameters(0)   # DDP semantics: every replica starts from rank 0's weights and buffers
        # The SGD(momentum) update is fused into the native backward, bucket by bucket (crimac_set_optimizer): a closed
        # bucket is updated on the communication stream while the rest of backward still runs.  Not with exchange="nccl"
        # (the all-reduce only happens after backward) and not for holders without the native plumbing (tests).
        self.fused_optimizer = bool(fused_optimizer) and self.exchange != "nccl" and dev.type == "cuda" and hasattr(model, "_set_native_opt")
        if self.fused_optimizer:
            if getattr(model, "_grad_arena", None) is None or model._grad_arena.numel() != total:
                self._arena_full = torch.zeros((total + 1023) // 1024 * 1024, dtype=torch.float32, device=dev)
                model._grad_arena = self._arena_full[:total]
            self._bind_optimizer()

    def _bind_optimizer(self):
        """(Re-)bind the fused optimizer of the model's native train contexts to the current learning rate."""
        opt = _engine._OptConfig()
        opt.params, opt.momentum = self.flat_params.data_ptr(), self.flat_momentum.data_ptr()
        opt.grads, opt.n = self.model._grad_arena.data_ptr(), self.flat_params.numel()
        opt.lr, opt.momentum_coef, opt.gscale = self.lr, self.momentum, 1.0 / self.world
        self.model._set_native_opt(opt)

    def broadcast_parameters(self, src=0):
        """Make every replica start from rank `src`'s weights and BN buffers."""
        if self.world > 1:
            dist.broadcast(self.flat_params, src)
            for b in self.model.buffers():
                dist.broadcast(b, src)

    def fit_host(self, batches):
        """Runs one optimisation step per (x, labels) pair of HOST tensors (pin them for full speed) and yields each
        step's loss as a 0-dim device tensor.  The host->device copy of batch i+1 runs on a copy stream while batch i
        is computed (two device buffers), which is what a DataLoader-fed loop (pipeline.py:161-178) needs to keep the
        GPU busy: the reference copies synchronously (`.to(device)`, pipeline.py:163-164)."""
        dev = self.flat_params.device
        copy_stream = torch.cuda.Stream(device=dev)
        main = torch.cuda.current_stream(dev)
        bufs, ready, freed = [None, None], [torch.cuda.Event(), torch.cuda.Event()], [None, None]

        def upload(slot, xh, yh):
            if bufs[slot] is None or bufs[slot][0].shape != xh.shape:
                bufs[slot] = (torch.empty(xh.shape, dtype=torch.float32, device=dev),
                              torch.empty(yh.shape, dtype=torch.int64, device=dev))
            with torch.cuda.stream(copy_stream):
                if freed[slot] is not None:
                    copy_stream.wait_event(freed[slot])   # the step that last read this buffer has finished
                bufs[slot][0].copy_(xh, non_blocking=True)
                bufs[slot][1].copy_(yh, non_blocking=True)
                ready[slot].record(copy_stream)

        it = iter(batches)
        nxt = next(it, None)
        if nxt is None:
            return
        upload(0, *nxt)
        i = 0
        while nxt is not None:
            slot = i & 1
            nxt = next(it, None)
            if nxt is not None:
                upload(slot ^ 1, *nxt)
            main.wait_event(ready[slot])
            loss = self.step(bufs[slot][0], bufs[slot][1])
            freed[slot] = torch.cuda.Event()
            freed[slot].record(main)
            yield loss
            i += 1

    def fit_survey(self, feeder, steps):
        """Runs `steps` optimisation steps on batches drawn on the device by a train_patches.SurveyPatchFeeder (the
        device-side replacement of the reference's DataLoader + Dataset.__getitem__, pipeline.py:161-178) and yields
        each step's loss as a 0-dim device tensor.  Everything is stream-ordered: no host synchronisation per step."""
        for _ in range(int(steps)):
            x, labels = feeder.next_batch()
            yield self.step(x, labels)

    def _launch_fwd_bwd(self, x, labels):
        """Forward + loss + backward: library kernels only (this is what a CUDA graph captures)."""
        return self.model.train_step_fused(x, labels, self.class_weight)

    def _launch_update(self):
        """Optimizer (+ the NCCL gradient exchange in the "nccl" A/B mode; with the peer exchange the gradients were
        already all-reduced inside backward).  A NCCL all-reduce is never captured in a graph (a captured collective
        keeps communicator resources alive and made process-group teardown hang in our runs)."""
        if self.fused_optimizer:
            return                                        # applied inside backward, bucket by bucket
        grads = self.model._grad_arena
        if self.exchange == "nccl":
            gscale = reduce_gradients(grads, self.world)  # NCCL over NVLink / NVSwitch
        else:
            gscale = 1.0 / self.world                     # peer exchange left the SUM over replicas in every arena
        _engine.sgd_step(self.flat_params, self.flat_momentum, grads, self.lr, self.momentum, gscale)

    def _launch_step(self, x, labels):
        loss = self._launch_fwd_bwd(x, labels)
        self._launch_update()
        return loss

    def _advance(self):
        # parameters and BatchNorm buffers were just written through raw pointers (fused SGD kernel, bn_finalize, also
        # inside a CUDA-graph replay): tell the model so that its eval path re-packs weights and re-folds BatchNorm
        self.model._native_mutated()
        self.iteration += 1
        if self.lr_step > 0 and self.iteration % self.lr_step == 0:
            self.lr *= self.lr_reduction  # ExponentialLR stepped every lr_step iterations (pipeline.py:157,188-189)
            if self.fused_optimizer:
                self._bind_optimizer()    # the learning rate is a launch argument of the fused update (graph re-captured)

    def step(self, x, labels):
        """One optimisation step on device tensors; returns the replica's loss as a 0-dim device tensor."""
        # the captured graph bakes in device addresses: re-capture if any parameter / buffer storage was replaced
        ptrs = hash(tuple(t.data_ptr() for t in self.model._state_tensors()))
        key = (tuple(x.shape), tuple(labels.shape), x.device, self.lr, ptrs)
        shape_key = key[:3]
        if not self.use_cuda_graph or shape_key not in self._eager_done:
            loss = self._launch_step(x, labels)           # first step of a shape: eager (lazy initialisation inside)
            self._eager_done.add(shape_key)
            self._advance()
            return loss
        if self._graph is None or self._graph_key != key:
            try:
                self._sx = torch.empty_like(x, dtype=torch.float32)
                self._sy = torch.empty_like(labels, dtype=torch.int64)
                torch.cuda.synchronize(x.device)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):                 # recorded, not executed
                    self._sloss = self._launch_fwd_bwd(self._sx, self._sy)
                    if self.exchange != "nccl":
                        self._launch_update()             # the SGD kernel (and the peer exchange) ride in the same graph
                self._graph, self._graph_key = g, key
            except Exception as exc:                      # e.g. a collective that cannot be captured: stay eager
                import warnings
                warnings.warn(f"CUDA graph capture of the train step failed ({exc!r}); using eager launches")
                self.use_cuda_graph = False
                self._graph = None
                loss = self._launch_step(x, labels)
                self._advance()
                return loss
        self._sx.copy_(x)
        self._sy.copy_(labels)
        self._graph.replay()
        if self.exchange == "nccl":
            self._launch_update()
        self._advance()
        return self._sloss.clone()
