"""Trainer step tail of the reference (train.py:86-88, 198-210) for the B200 path:

    loss.mean().backward(); clip_grad_value_(params, 5.0); Adam(amsgrad=True, weight_decay=1e-5).step()

re-laid out for one GPU per process: parameters and gradients live in ONE flat fp32 buffer each
(the module's ``nn.Parameter``s become views, so ``state_dict`` / checkpoints are unchanged), the
clip + Adam update is one kernel launch over that buffer (``mvc_clip_adam_step``) and, under data
parallelism, the gradient exchange is an NCCL all-reduce of the flat gradient buffer over
NVLink/NVSwitch -- in buckets, launched on a communication stream as the backward
pass finishes each group of gradients (``GradBuckets``) -- with the 1/world averaging folded into
the update kernel.

``FlatClipAdam`` is a ``torch.optim.Optimizer`` (one param group): ``ReduceLROnPlateau(optimizer)``
(train.py:90-97) drives its ``lr`` like any other optimizer's, and ``state_dict`` round-trips.
Parameters that never receive a gradient (``AVCaptioningDual.output_fc``, captioning.py:185) are
left untouched, as ``torch.optim.Adam`` does for ``grad is None``.
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Tuple

import ctypes as C

import torch

from . import functional as Fn


def C_void(addr: int):
    import ctypes
    return ctypes.c_void_p(int(addr))


class FlatClipAdam(torch.optim.Optimizer):
    def __init__(self, params: Iterable[torch.nn.Parameter], lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-5,
                 clip_value: float = 5.0, process_group=None, world_size: Optional[int] = None, amsgrad: bool = True,
                 fused_comm: Optional[bool] = None):
        if not amsgrad:
            raise NotImplementedError("FlatClipAdam implements the reference's Adam(amsgrad=True) (train.py:86-88)")
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, clip_value=clip_value)
        super().__init__([p for p in params if p.requires_grad], defaults)
        if len(self.param_groups) != 1:
            raise ValueError("FlatClipAdam takes one parameter group (the reference trains model.parameters())")
        self.params: List[torch.nn.Parameter] = self.param_groups[0]["params"]
        self.group = process_group
        self.world = world_size
        self.step_count = 0
        self.flat_p = self.flat_g = self.m = self.v = self.vmax = None
        self._live: Optional[List[torch.nn.Parameter]] = None
        self._live_set = set()
        self._views: List[torch.Tensor] = []
        self._offsets: List[int] = []
        self._n_flat = 0            # elements of the flat buffers in use (every parameter padded to a multiple of 4)
        self._arena = None
        self._pending = []          # async all-reduce work handles of this step (bucketed mode)
        self._dev_state = None      # [step count, lr] on the device: set by GraphedTrainStep (CUDA-graph replay)
        # fused_comm: exchange gradients and update parameters in ONE kernel over NVSwitch multicast
        # (mvc_clip_adam_multimem: in-switch reduce-scatter + sharded clip/Adam + multicast all-gather) instead of an NCCL
        # all-reduce followed by a full-size update.  None = use it when world > 1 and the fabric supports multicast.
        self.fused_comm = fused_comm
        self._mc = None             # (symmetric-memory handle of flat_p, handle of flat_g, lo, hi) when active
        # reduce-scatter half of the fused exchange: "peer" = the owner of a slice loads the N replicas itself
        # (mvc_clip_adam_p2p_multimem), "switch" = multimem.ld_reduce (mvc_clip_adam_multimem)
        import os
        # (measured at 2 GPUs: 253.0 k samples/s with peer loads, 256.6 k with the in-switch reduction -- the latter stays
        # the default although it moves a third more NVLink bytes)
        self._reduce = os.environ.get("MVC_B200_FUSED_REDUCE", "switch")
        if self._reduce not in ("peer", "switch"):
            raise ValueError("MVC_B200_FUSED_REDUCE must be 'peer' or 'switch'")

    # convenience mirrors of the single param group (kept in sync with lr schedulers)
    @property
    def lr(self):
        return self.param_groups[0]["lr"]

    @lr.setter
    def lr(self, value):
        self.param_groups[0]["lr"] = value

    # ---- lazy flattening: after the first backward we know which parameters get gradients
    def _flatten(self):
        live = [p for p in self.params if p.grad is not None]
        if not live:
            raise RuntimeError("FlatClipAdam.step() before any backward()")
        dev = live[0].device
        old = None
        if self.flat_p is not None:       # a parameter received its first gradient later: re-flatten, keep the moments
            old = {id(p): (o, p.numel()) for p, o in zip(self._live, self._offsets)}
            old_m, old_v, old_vmax = self.m, self.v, self.vmax
        # every parameter starts on a 16-byte boundary of the flat buffers (the kernels' float4 paths -- weight packing,
        # casts, GEMM epilogues writing gradients -- fall back to scalar code on a misaligned view: one odd-sized tensor,
        # e.g. out.bias [3201], would push every later parameter off); the <= 3 pad elements in front of a parameter stay
        # zero in every buffer (zero gradient, zero weight: the update leaves them at zero)
        al = lambda x: (x + 3) // 4 * 4
        n = sum(al(p.numel()) for p in live)
        flat_p, flat_g = self._alloc_flat(n, dev)
        flat_p.zero_(); flat_g.zero_()
        m, v, vmax = (torch.zeros(flat_p.numel(), device=dev, dtype=torch.float32) for _ in range(3))
        off, offsets = 0, []
        for p in live:
            k = p.numel()
            flat_p[off:off + k].copy_(p.data.reshape(-1))
            flat_g[off:off + k].copy_(p.grad.reshape(-1))
            if old is not None and id(p) in old:
                o = old[id(p)][0]
                m[off:off + k].copy_(old_m[o:o + k]); v[off:off + k].copy_(old_v[o:o + k])
                vmax[off:off + k].copy_(old_vmax[o:o + k])
            p.data = flat_p[off:off + k].view_as(p)
            p.grad = flat_g[off:off + k].view_as(p)
            offsets.append(off)
            off += al(k)
        self._n_flat = off
        self.flat_p, self.flat_g, self.m, self.v, self.vmax = flat_p, flat_g, m, v, vmax
        self._live, self._offsets = live, offsets
        self._live_set = {id(p) for p in live}
        self._views = [p.grad for p in live]
        if dev.type == "cuda":
            # backward kernels write gradients straight into these views (functional.GradArena)
            self._arena = Fn.GradArena(live, self._views)
            Fn.register_grad_arena(self._arena)

    def _alloc_flat(self, n: int, dev):
        """Flat parameter / gradient buffers.  Under data parallelism on an NVSwitch fabric they are allocated as
        SYMMETRIC memory (same size on every rank) and bound to a multicast object, so that the fused exchange + update
        kernel can address all replicas at once; otherwise plain device memory."""
        self._mc = None
        world = self.world or 1
        want = self.fused_comm if self.fused_comm is not None else (world > 1)
        if want and world > 1 and dev.type == "cuda":
            try:
                import torch.distributed as dist
                import torch.distributed._symmetric_memory as symm_mem
                group = self.group if self.group is not None else dist.group.WORLD
                n_pad = (n + 4 * world - 1) // (4 * world) * (4 * world)          # every rank owns a float4-aligned slice
                flat_p = symm_mem.empty(n_pad, dtype=torch.float32, device=dev)
                flat_g = symm_mem.empty(n_pad, dtype=torch.float32, device=dev)
                hp, hg = symm_mem.rendezvous(flat_p, group), symm_mem.rendezvous(flat_g, group)
                if hp.multicast_ptr and hg.multicast_ptr:
                    flat_p.zero_(); flat_g.zero_()
                    per = n_pad // world
                    rank = dist.get_rank(group)
                    self._mc = (hp, hg, rank * per, (rank + 1) * per)
                    return flat_p, flat_g
                if self.fused_comm:
                    raise RuntimeError("no multicast support on this fabric")
            except Exception as e:
                if self.fused_comm:
                    raise RuntimeError(f"FlatClipAdam(fused_comm=True): symmetric-memory / multicast setup failed: {e}")
        return (torch.empty(n, device=dev, dtype=torch.float32), torch.empty(n, device=dev, dtype=torch.float32))

    def zero_grad(self, set_to_none: bool = True):
        """Before the first step: plain ``grad = None``.  Afterwards gradients live in the flat buffer; on CUDA the
        backward kernels OVERWRITE their arena views (and autograd adopts them because ``grad is None``), so no
        zero-fill is needed; with ``set_to_none=False`` (or on CPU, in the plumbing tests) the buffer is cleared and
        autograd accumulates into it."""
        if self._arena is not None:
            self._arena.new_step()
        if self.flat_g is None or (set_to_none and self.flat_g.device.type == "cuda"):
            for p in self.params:
                p.grad = None
        else:
            self.flat_g.zero_()
            for p, v in zip(self._live, self._views):
                p.grad = v

    def _sync_views(self) -> List[Tuple[int, int]]:
        """Make every live parameter's .grad its arena view (copy in anything autograd allocated itself) and return
        the [lo, hi) element ranges of the flat buffer that hold a gradient this step.  A parameter whose grad is
        None is skipped by torch.optim.Adam (no weight decay, no moment update): its range is left out."""
        if self._live is None or any(p.grad is not None and id(p) not in self._live_set for p in self.params):
            self._flatten()
        ranges: List[Tuple[int, int]] = []
        ends = self._offsets[1:] + [self._n_flat]            # a parameter's range runs up to the next one (pads included)
        for p, v, o, e in zip(self._live, self._views, self._offsets, ends):
            g = p.grad
            if g is None:
                continue
            if g.data_ptr() != v.data_ptr():
                v.copy_(g)
                p.grad = v
            if ranges and ranges[-1][1] == o:
                ranges[-1] = (ranges[-1][0], e)
            else:
                ranges.append((o, e))
        return ranges

    # ---- gradient exchange
    def all_reduce_grads(self):
        """NCCL all-reduce (sum) of the flat gradient buffer; averaging happens in step().  Buckets already
        exchanged by the backward-pass hook (GradBuckets) are only waited for."""
        import torch.distributed as dist
        self._sync_views()
        if self._mc is not None:
            return                      # the exchange happens inside step() (mvc_clip_adam_multimem)
        if self._pending:
            for w in self._pending:
                w.wait()
            self._pending = []
            return
        dist.all_reduce(self.flat_g, op=dist.ReduceOp.SUM, group=self.group)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        ranges = self._sync_views()
        if self.flat_p.device.type != "cuda":
            raise RuntimeError("FlatClipAdam.step: parameters must live on CUDA (the update is a CUDA kernel; "
                               "there is no CPU fallback)")
        for w in self._pending:
            w.wait()
        self._pending = []
        self.step_count += 1
        g = self.param_groups[0]
        scale = 1.0 / self.world if (self.world is not None and self.world > 1) else 1.0
        if self._mc is not None:
            self._fused_exchange_and_update(ranges, g, scale)
            if self._arena is not None:
                self._arena.new_step()
            return loss
        for k, (lo, hi) in enumerate(ranges):
            if self._dev_state is not None:      # graph-capturable form: step count and lr live on the device
                from . import cabi
                cabi.check(cabi.lib().mvc_clip_adam_step_dev(
                    cabi.ptr(self.flat_p[lo:hi]), cabi.ptr(self.flat_g[lo:hi]), cabi.ptr(self.m[lo:hi]), cabi.ptr(self.v[lo:hi]),
                    cabi.ptr(self.vmax[lo:hi]), hi - lo, cabi.ptr(self._dev_state), int(k == 0), g["betas"][0], g["betas"][1],
                    g["eps"], g["weight_decay"], g["clip_value"], scale, cabi.stream_ptr()), "mvc_clip_adam_step_dev")
                continue
            Fn.clip_adam_step(self.flat_p[lo:hi], self.flat_g[lo:hi], self.m[lo:hi], self.v[lo:hi], self.vmax[lo:hi],
                              lr=g["lr"], betas=g["betas"], eps=g["eps"], weight_decay=g["weight_decay"],
                              clip_value=g["clip_value"], step=self.step_count, grad_scale=scale)
        if self._arena is not None:
            self._arena.new_step()
        return loss

    def _fused_exchange_and_update(self, ranges, g, scale):
        """barrier (every rank's gradients are complete) -> ONE kernel: in-switch sum of the replicas' gradients for the
        slice this rank owns, clip + Adam on that slice, multicast store of the new parameters into every replica ->
        barrier (every slice has landed everywhere)."""
        from . import cabi
        hp, hg, lo, hi = self._mc
        if ranges != [(0, self._n_flat)]:
            raise RuntimeError("FlatClipAdam(fused_comm): every live parameter must receive a gradient each step")
        if self._dev_state is None:
            self._dev_state = torch.tensor([float(self.step_count - 1), float(g["lr"])], device=self.flat_p.device)
            self._lr_on_dev = float(g["lr"])
        elif getattr(self, "_lr_on_dev", None) != float(g["lr"]) and not torch.cuda.is_current_stream_capturing():
            self._dev_state[1] = float(g["lr"])
            self._lr_on_dev = float(g["lr"])
        hg.barrier(channel=0)
        if self._reduce == "switch":
            cabi.check(cabi.lib().mvc_clip_adam_multimem(
                cabi.ptr(self.flat_p), C_void(hp.multicast_ptr), C_void(hg.multicast_ptr), cabi.ptr(self.m), cabi.ptr(self.v),
                cabi.ptr(self.vmax), lo, hi, cabi.ptr(self._dev_state), 1, g["betas"][0], g["betas"][1], g["eps"],
                g["weight_decay"], g["clip_value"], scale, cabi.stream_ptr()), "mvc_clip_adam_multimem")
        else:
            # reduce-scatter by peer loads (the owner reads the N replicas itself): fewer NVLink bytes than the in-switch
            # reduction, which fetches the requester's own replica over the link as well
            ptrs = [int(x) for x in hg.buffer_ptrs]
            arr = (C.c_void_p * len(ptrs))(*ptrs)
            cabi.check(cabi.lib().mvc_clip_adam_p2p_multimem(
                cabi.ptr(self.flat_p), C_void(hp.multicast_ptr), arr, len(ptrs), cabi.ptr(self.m), cabi.ptr(self.v),
                cabi.ptr(self.vmax), lo, hi, cabi.ptr(self._dev_state), 1, g["betas"][0], g["betas"][1], g["eps"],
                g["weight_decay"], g["clip_value"], scale, cabi.stream_ptr()), "mvc_clip_adam_p2p_multimem")
        hp.barrier(channel=1)

    # ---- checkpointing: the flat moments, addressed by parameter order
    def state_dict(self):
        sd = {"step": self.step_count, "param_group": {k: v for k, v in self.param_groups[0].items() if k != "params"}}
        if self.flat_p is not None:
            # moments concatenated per live parameter, WITHOUT the alignment pads: independent of the buffer layout
            pick = lambda buf: torch.cat([buf[o:o + p.numel()] for p, o in zip(self._live, self._offsets)])
            sd.update(exp_avg=pick(self.m), exp_avg_sq=pick(self.v), max_exp_avg_sq=pick(self.vmax),
                      live=[i for i, p in enumerate(self.params) if id(p) in self._live_set])
        return sd

    def load_state_dict(self, sd):
        self.step_count = int(sd["step"])
        self.param_groups[0].update(sd["param_group"])
        if "exp_avg" in sd:
            live = [self.params[i] for i in sd["live"]]
            for p in live:
                if p.grad is None:
                    p.grad = torch.zeros_like(p)
            keep = {id(p) for p in live}
            for p in self.params:
                if id(p) not in keep:
                    p.grad = None
            self.flat_p = None
            self._flatten()
            src = 0
            for p, o in zip(self._live, self._offsets):
                k = p.numel()
                self.m[o:o + k].copy_(sd["exp_avg"][src:src + k]); self.v[o:o + k].copy_(sd["exp_avg_sq"][src:src + k])
                self.vmax[o:o + k].copy_(sd["max_exp_avg_sq"][src:src + k])
                src += k


class GraphedTrainStep:
    """One optimiser step of the reference's train loop (train.py:186-210) -- forward, ModalityWiseReconstructionLoss,
    backward, gradient all-reduce (world > 1), clip + Adam(amsgrad) -- recorded ONCE as a CUDA graph and replayed:

        step = GraphedTrainStep(model, loss_fn, FlatClipAdam(model.parameters(), ...), example_batch)
        for audio, visual, captions, _ in feeder:
            terms = step(audio, visual, captions)      # (loss, ce, entropy, a_rec, v_rec): device scalars of THIS step

    The eager step issues ~80 kernel launches, a dozen memsets and ~40 ctypes / autograd calls from Python: ~1.05 ms of
    host time per step against ~1.3 ms of GPU time at the MSVD shape; a replay is one launch.  Requirements (checked):
    teacher_forcing_ratio == 1 (no per-step host RNG decision), fixed batch shape; the inputs of each call are copied
    into the graph's static input tensors on the device (one small copy kernel per tensor).  The step count and the
    learning rate live in device memory (mvc_clip_adam_step_dev): lr schedulers keep working through
    ``optimizer.param_groups[0]["lr"]``, which is pushed to the device whenever it changes.

    ``slots=K`` (K > 1) records K graphs over K sets of static input tensors (``input_slots``) that share one memory
    pool.  A caller that fills those tensors itself -- ``ShardFeeder(..., device_slots=step.input_slots)`` uploads every
    batch straight into them -- and passes them back gets the step WITHOUT the device-to-device input copy (22 us of a
    1.1 ms step at the MSVD shape); any other tensors are copied into slot 0 as before.
    """

    def __init__(self, model, loss_fn, optimizer: FlatClipAdam, example_batch, warmup: int = 3, slots: int = 1):
        audio, visual, captions = example_batch[:3]
        if getattr(model, "teacher_forcing_ratio", 1.0) != 1.0:
            raise ValueError("GraphedTrainStep needs teacher_forcing_ratio == 1.0 (the per-step RNG draw is a host decision)")
        dev = audio.device
        if dev.type != "cuda":
            raise RuntimeError("GraphedTrainStep: inputs must live on CUDA")
        self.model, self.loss_fn, self.opt = model, loss_fn, optimizer
        self._ins = [tuple(torch.empty_like(t) for t in (audio, visual, captions)) for _ in range(max(1, int(slots)))]
        for ins in self._ins:
            for dst, src in zip(ins, (audio, visual, captions)):
                dst.copy_(src)
        self._in = self._ins[0]
        self._slot_of = {ins[0].data_ptr(): k for k, ins in enumerate(self._ins)}
        self._lr = float(optimizer.param_groups[0]["lr"])
        multi = optimizer.world is not None and optimizer.world > 1

        def eager(ins=None):
            ins = self._in if ins is None else ins
            optimizer.zero_grad()
            out, ar, vr = model(*ins)
            terms = loss_fn(out, ins[2], ins[0], ar, ins[1], vr)
            loss = terms[0]
            (loss if loss.numel() == 1 else loss.mean()).backward()       # mean() of one element: two kernels for nothing
            if multi:
                optimizer.all_reduce_grads()
            optimizer.step()
            return terms

        # warm-up on the capture stream: lazy initialisations (side stream, scratch buffers, occupancy queries, the flat
        # parameter / gradient buffers and the gradient arena) must all have happened before the capture starts
        # the capture stream outranks the library's side streams (priority 0): when both have blocks waiting -- operand
        # preparation next to the launch of a persistent kernel -- the critical chain gets the SMs first
        stream = torch.cuda.Stream(device=dev, priority=-1)
        stream.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(stream):
            for _ in range(max(2, warmup)):
                eager()
        stream.synchronize()
        optimizer._dev_state = torch.tensor([float(optimizer.step_count), self._lr], device=dev, dtype=torch.float32)
        self._graphs, self._terms_of = [], []
        for k, ins in enumerate(self._ins):
            g = torch.cuda.CUDAGraph()
            optimizer.zero_grad()
            kw = {"pool": self._graphs[0].pool()} if k else {}       # replays are serial: one pool serves every slot
            with torch.cuda.graph(g, stream=stream, capture_error_mode="thread_local", **kw):
                terms = eager(ins)
            self._graphs.append(g)
            self._terms_of.append(tuple(terms))
            optimizer.step_count -= 1             # the capture pass recorded, but did not execute, one update
        self.graph, self._terms = self._graphs[0], self._terms_of[0]

    @property
    def input_slots(self):
        """The static (audio, visual, captions) tensors of every recorded graph: fill one and pass it to __call__."""
        return list(self._ins)

    def __call__(self, audio, visual, captions):
        lr = float(self.opt.param_groups[0]["lr"])
        if lr != self._lr:
            self.opt._dev_state[1] = lr
            self._lr = lr
        k = self._slot_of.get(audio.data_ptr(), -1)
        if k >= 0 and visual.data_ptr() == self._ins[k][1].data_ptr() and captions.data_ptr() == self._ins[k][2].data_ptr():
            self._graphs[k].replay()              # the caller filled the graph's own input tensors: no copy
        else:
            k = 0
            self._in[0].copy_(audio, non_blocking=True)
            self._in[1].copy_(visual, non_blocking=True)
            self._in[2].copy_(captions, non_blocking=True)
            self.graph.replay()
        self.opt.step_count += 1
        return self._terms_of[k]


def shard_batch(audio, visual, captions, rank: int, world: int):
    """Contiguous split of the batch dimension (features are batch-first, captions time-first);
    SURVEY.md §8e.  Inference sharding needs no communication."""
    B = audio.shape[0]
    per = (B + world - 1) // world
    lo, hi = min(B, rank * per), min(B, (rank + 1) * per)
    return audio[lo:hi], visual[lo:hi], None if captions is None else captions[:, lo:hi]
