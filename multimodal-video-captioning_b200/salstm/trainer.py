"""Trainer step tail of the reference (train.py:86-88, 198-210) for the B200 path:

    loss.mean().backward(); clip_grad_value_(params, 5.0); Adam(amsgrad=True, weight_decay=1e-5).step()

re-laid out for one GPU per process: parameters and gradients live in ONE flat fp32 buffer each
(the module's ``nn.Parameter``s become views, so ``state_dict`` / checkpoints are unchanged), the
clip + Adam update is one kernel launch over that buffer (``mvc_clip_adam_step``) and, under data
parallelism, the gradient exchange is one NCCL all-reduce of the flat gradient buffer over
NVLink/NVSwitch with the 1/world averaging folded into the same kernel.

Parameters that never receive a gradient (``AVCaptioningDual.output_fc``, captioning.py:185) are
left untouched, as ``torch.optim.Adam`` does for ``grad is None``.
"""
from __future__ import annotations

from typing import Iterable, Optional

import torch

from . import functional as Fn


class FlatClipAdam:
    def __init__(self, params: Iterable[torch.nn.Parameter], lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-5,
                 clip_value: float = 5.0, process_group=None, world_size: Optional[int] = None):
        self.params = [p for p in params if p.requires_grad]
        self.lr, self.betas, self.eps, self.weight_decay, self.clip_value = lr, betas, eps, weight_decay, clip_value
        self.group = process_group
        self.world = world_size
        self.step_count = 0
        self.flat_p = self.flat_g = self.m = self.v = self.vmax = None
        self._live = None

    # ---- lazy flattening: after the first backward we know which parameters get gradients
    def _flatten(self):
        live = [p for p in self.params if p.grad is not None]
        if not live:
            raise RuntimeError("FlatClipAdam.step() before any backward()")
        dev = live[0].device
        n = sum(p.numel() for p in live)
        self.flat_p = torch.empty(n, device=dev, dtype=torch.float32)
        self.flat_g = torch.empty(n, device=dev, dtype=torch.float32)
        off = 0
        for p in live:
            k = p.numel()
            self.flat_p[off:off + k].copy_(p.data.reshape(-1))
            self.flat_g[off:off + k].copy_(p.grad.reshape(-1))
            p.data = self.flat_p[off:off + k].view_as(p)
            p.grad = self.flat_g[off:off + k].view_as(p)
            off += k
        self.m, self.v, self.vmax = (torch.zeros_like(self.flat_p) for _ in range(3))
        self._live = live
        self._views = [p.grad for p in live]
        if dev.type == "cuda":
            # backward kernels write gradients straight into these views (functional._GRAD_ARENA)
            Fn.register_grad_arena({p.data_ptr(): g for p, g in zip(live, self._views)})

    def zero_grad(self):
        """Before the first step: plain ``grad = None``.  Afterwards gradients live in the flat buffer; on CUDA the
        backward kernels OVERWRITE their arena views (and autograd adopts them because ``grad is None``), so no
        zero-fill is needed; elsewhere (CPU plumbing tests) the buffer is cleared and autograd accumulates."""
        if self.flat_g is None or self.flat_g.device.type == "cuda":
            for p in self.params:
                p.grad = None
        else:
            self.flat_g.zero_()

    def _adopt(self):
        """Make sure every live parameter's .grad IS its arena view (copy in anything autograd allocated itself)."""
        for p, v in zip(self._live, self._views):
            g = p.grad
            if g is None:
                v.zero_()
            elif g.data_ptr() != v.data_ptr():
                v.copy_(g)
            p.grad = v

    def all_reduce_grads(self):
        """One NCCL all-reduce (sum) over the flat gradient buffer; averaging happens in step()."""
        import torch.distributed as dist
        if self.flat_g is None:
            self._flatten()
        else:
            self._adopt()
        dist.all_reduce(self.flat_g, op=dist.ReduceOp.SUM, group=self.group)

    def step(self):
        if self.flat_g is None:
            self._flatten()
        else:
            self._adopt()
        if self.flat_p.device.type != "cuda":
            raise RuntimeError("FlatClipAdam.step: parameters must live on CUDA (the update is a CUDA kernel; "
                               "there is no CPU fallback)")
        self.step_count += 1
        scale = 1.0
        if self.world is not None and self.world > 1:
            scale = 1.0 / self.world
        Fn.clip_adam_step(self.flat_p, self.flat_g, self.m, self.v, self.vmax, lr=self.lr, betas=self.betas,
                          eps=self.eps, weight_decay=self.weight_decay, clip_value=self.clip_value,
                          step=self.step_count, grad_scale=scale)


def shard_batch(audio, visual, captions, rank: int, world: int):
    """Contiguous split of the batch dimension (features are batch-first, captions time-first);
    SURVEY.md §8e.  Inference sharding needs no communication."""
    B = audio.shape[0]
    per = (B + world - 1) // world
    lo, hi = min(B, rank * per), min(B, (rank + 1) * per)
    return audio[lo:hi], visual[lo:hi], None if captions is None else captions[:, lo:hi]
