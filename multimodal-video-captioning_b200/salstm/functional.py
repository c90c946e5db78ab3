"""torch.autograd glue over the C ABI: owns save-for-backward, nothing else.

Every Function below is a thin wrapper: allocate outputs / workspace with
torch, make ONE call into libmvc_b200 for the forward and ONE for the backward.
All arithmetic happens in the CUDA kernels.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import torch

from . import cabi

DEC_PARAM_ORDER = ("embedding.weight", "attention.W.weight", "attention.U.weight", "attention.b",
                   "attention.w.weight", "rnn.weight_ih_l0", "rnn.weight_hh_l0", "rnn.bias_ih_l0",
                   "rnn.bias_hh_l0", "out.weight", "out.bias")
REC_PARAM_ORDER = ("rnn.weight_ih_l0", "rnn.weight_hh_l0", "rnn.bias_ih_l0", "rnn.bias_hh_l0",
                   "attention.W.weight", "attention.U.weight", "attention.b", "attention.w.weight")


# Gradient arena (salstm/trainer.py::FlatClipAdam): parameter storage address -> view of the flat gradient
# buffer.  When a parameter is registered here, the backward kernels write its gradient straight into the
# arena view and autograd (with .grad == None) adopts that view: no zero-fill, no accumulate pass.
_GRAD_ARENA: dict = {}


def register_grad_arena(mapping: dict) -> None:
    _GRAD_ARENA.clear()
    _GRAD_ARENA.update(mapping)


def _grad_like(t: torch.Tensor) -> torch.Tensor:
    v = _GRAD_ARENA.get(t.data_ptr())
    if v is not None and v.shape == t.shape and v.device == t.device:
        return v.detach()          # a fresh alias: autograd may adopt it as .grad without cloning
    return torch.empty_like(t)


def _f32c(t: Optional[torch.Tensor], name: str) -> Optional[torch.Tensor]:
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError(f"mvc_b200: `{name}` lives on {t.device}; this path runs on CUDA only (no CPU fallback)")
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _feats(audio, visual, prec):
    """Decoder feature inputs: fp32 as the reference's loader hands them (get_loader.py:242-268), or -- bf16 compute
    path only -- bf16 tensors (pre-packed feature shards), passed to the library uncast.
    Returns (audio, visual, input_format)."""
    given = [t for t in (audio, visual) if t is not None]
    if given and prec == cabi.MVC_BF16 and all(t.dtype == torch.bfloat16 for t in given):
        for t in given:
            if not t.is_cuda:
                raise RuntimeError(f"mvc_b200: features live on {t.device}; this path runs on CUDA only (no CPU fallback)")
            if t.shape[-1] % 8:
                raise ValueError("bf16 feature shards need feature widths that are multiples of 8")
        return (None if audio is None else audio.contiguous(), None if visual is None else visual.contiguous(),
                cabi.MVC_INPUT_BF16)
    return _f32c(audio, "audio"), _f32c(visual, "visual"), cabi.MVC_INPUT_F32


class _input_format:
    """Declares the feature format for the library calls inside the block (thread-local in the library)."""

    def __init__(self, fmt):
        self.fmt = fmt

    def __enter__(self):
        if self.fmt != cabi.MVC_INPUT_F32:
            cabi.check(cabi.lib().mvc_set_input_format(self.fmt), "mvc_set_input_format")

    def __exit__(self, *exc):
        if self.fmt != cabi.MVC_INPUT_F32:
            cabi.lib().mvc_set_input_format(cabi.MVC_INPUT_F32)
        return False


def _dec_structs(dims, params: Sequence[torch.Tensor]):
    d = cabi.DecoderDims(*dims)
    p = cabi.DecoderParams(*[C.c_void_p(t.data_ptr()) for t in params])
    return d, p


def teacher_flags(captions, max_len: int, ratio: float) -> List[bool]:
    """One ``torch.rand(1) < ratio`` draw per loop step from the global CPU RNG iff
    captions are given -- the same stream of draws as features_captioning.py:113-116."""
    if captions is None:
        return [False] * (max_len - 1)
    n = max_len - 1
    if 0 < n <= 64:
        # one draw of n values consumes the CPU generator exactly like n draws of one value (same values, same
        # final state; checked in tests/test_oracle_golden.py) at a fraction of the host cost
        return (torch.rand(n) < ratio).tolist()
    return [bool(torch.rand(1) < ratio) for _ in range(1, max_len)]


class DecoderFn(torch.autograd.Function):
    """FeaturesCaptioning.decode (features_captioning.py:91-129) as one autograd node."""

    @staticmethod
    def forward(ctx, dims, flags, audio, visual, captions, *params):
        lib = cabi.lib()
        B, T, F, H, E, A, V, L, prec = dims
        params = [_f32c(p.detach(), "parameter") for p in params]
        audio, visual, fmt = _feats(audio, visual, prec)
        dev = params[0].device
        Fa = 0 if audio is None else audio.shape[-1]
        Fv = 0 if visual is None else visual.shape[-1]
        if captions is not None:
            captions = captions.to(device=dev, dtype=torch.int64).contiguous()
        d, p = _dec_structs(dims, params)
        out = torch.empty(L, B, V, device=dev, dtype=torch.float32)
        hid = torch.empty(L, B, H, device=dev, dtype=torch.float32)
        tokens = torch.empty(max(L - 1, 1), B, device=dev, dtype=torch.int64)
        nbytes = lib.mvc_decoder_fwd_workspace_bytes(C.byref(d), 1)
        ws = cabi.workspace(nbytes, dev)
        fl = (C.c_uint8 * max(L - 1, 1))(*[1 if f else 0 for f in flags])
        with _input_format(fmt):
            cabi.check(lib.mvc_decoder_forward(C.byref(d), C.byref(p), cabi.ptr(audio), Fa, cabi.ptr(visual), Fv,
                                               cabi.ptr(captions), fl, cabi.ptr(out), cabi.ptr(hid), cabi.ptr(tokens),
                                               cabi.ptr(ws), nbytes, 1, cabi.stream_ptr()), "mvc_decoder_forward")
        ctx.dims = dims
        ctx.save_for_backward(out, tokens, ws, *params)
        return out, hid.unsqueeze(1)

    @staticmethod
    def backward(ctx, dout, dhid):
        lib = cabi.lib()
        out, tokens, ws, *params = ctx.saved_tensors
        dims = ctx.dims
        d, p = _dec_structs(dims, params)
        dout = None if dout is None else _f32c(dout, "grad")
        dhid = None if dhid is None else _f32c(dhid, "grad")
        grads = [_grad_like(t) for t in params]
        g = cabi.DecoderGrads(*[C.c_void_p(t.data_ptr()) for t in grads])
        nbytes = lib.mvc_decoder_bwd_workspace_bytes(C.byref(d))
        bws = cabi.workspace(nbytes, out.device)
        cabi.check(lib.mvc_decoder_backward(C.byref(d), C.byref(p), cabi.ptr(out), cabi.ptr(dout), cabi.ptr(dhid),
                                            cabi.ptr(tokens), cabi.ptr(ws), C.byref(g), cabi.ptr(bws), nbytes,
                                            cabi.stream_ptr()), "mvc_decoder_backward")
        return (None, None, None, None, None, *grads)


def decoder_greedy(dims, audio, visual, params) -> torch.Tensor:
    """ids [B, L] int64 (column 0 = 0); captioning.py:138-141."""
    lib = cabi.lib()
    B, L = dims[0], dims[7]
    params = [_f32c(p.detach(), "parameter") for p in params]
    audio, visual, fmt = _feats(audio, visual, dims[8])
    dev = params[0].device
    d, p = _dec_structs(dims, params)
    ids = torch.empty(B, L, device=dev, dtype=torch.int64)
    nbytes = lib.mvc_decoder_greedy_workspace_bytes(C.byref(d))
    ws = cabi.workspace(nbytes, dev)
    with _input_format(fmt):
        cabi.check(lib.mvc_decoder_greedy(C.byref(d), C.byref(p), cabi.ptr(audio), 0 if audio is None else audio.shape[-1],
                                          cabi.ptr(visual), 0 if visual is None else visual.shape[-1], cabi.ptr(ids),
                                          cabi.ptr(ws), nbytes, cabi.stream_ptr()), "mvc_decoder_greedy")
    return ids


def decoder_beam(dims, audio, visual, params, width: int, alpha: float) -> torch.Tensor:
    """ids [B, L+2] int64 = SOS + L+1 tokens of the best beam; features_captioning.py:131-228."""
    lib = cabi.lib()
    B, L = dims[0], dims[7]
    params = [_f32c(p.detach(), "parameter") for p in params]
    audio, visual, fmt = _feats(audio, visual, dims[8])
    dev = params[0].device
    d, p = _dec_structs(dims, params)
    ids = torch.empty(B, L + 2, device=dev, dtype=torch.int64)
    nbytes = lib.mvc_decoder_beam_workspace_bytes(C.byref(d), int(width))
    ws = cabi.workspace(nbytes, dev)
    with _input_format(fmt):
        cabi.check(lib.mvc_decoder_beam(C.byref(d), C.byref(p), cabi.ptr(audio), 0 if audio is None else audio.shape[-1],
                                        cabi.ptr(visual), 0 if visual is None else visual.shape[-1], int(width),
                                        float(alpha), cabi.ptr(ids), cabi.ptr(ws), nbytes, cabi.stream_ptr()),
                   "mvc_decoder_beam")
    return ids


# --------------------------------------------------------------------------- reconstructors
def caption_mask(outputs: torch.Tensor, captions: Optional[torch.Tensor]) -> torch.Tensor:
    """build_caption_mask (reconstructor.py:197-206) -> uint8 [L,B]."""
    lib = cabi.lib()
    if captions is None:
        L, B, V = outputs.shape
        x = _f32c(outputs.detach(), "outputs")
        captions = torch.empty(L, B, device=x.device, dtype=torch.int64)
        cabi.check(lib.mvc_argmax_rows(cabi.ptr(x), None, L * B, V, cabi.ptr(captions), cabi.stream_ptr()), "argmax")
    captions = captions.to(device=outputs.device, dtype=torch.int64).contiguous()
    mask = torch.empty(captions.shape, device=captions.device, dtype=torch.uint8)
    cabi.check(lib.mvc_caption_mask(cabi.ptr(captions), captions.numel(), cabi.ptr(mask), cabi.stream_ptr()), "mask")
    return mask


def _rec_structs(dims, params):
    d = cabi.ReconDims(*dims)
    ptrs = [C.c_void_p(t.data_ptr()) for t in params] + [None] * (8 - len(params))
    return d, cabi.ReconParams(*ptrs)


class _ReconFn(torch.autograd.Function):
    KIND = "global"

    @staticmethod
    def _fns(lib, kind):
        return (getattr(lib, f"mvc_{kind}_recon_workspace_bytes"), getattr(lib, f"mvc_{kind}_recon_bwd_workspace_bytes"),
                getattr(lib, f"mvc_{kind}_recon_forward"), getattr(lib, f"mvc_{kind}_recon_backward"))

    @classmethod
    def _forward(cls, ctx, dims, hid, mask, params):
        lib = cabi.lib()
        ws_b, _, fwd, _ = cls._fns(lib, cls.KIND)
        B, L, H, Fr, A, T, prec = dims
        params = [_f32c(p.detach(), "parameter") for p in params]
        hid3 = _f32c(hid.detach(), "decoder_hiddens").reshape(L, B, H)
        d, p = _rec_structs(dims, params)
        n_out = L if cls.KIND == "global" else T
        rec = torch.empty(B, n_out, Fr, device=hid3.device, dtype=torch.float32)
        nbytes = ws_b(C.byref(d))
        ws = cabi.workspace(nbytes, hid3.device)
        cabi.check(fwd(C.byref(d), C.byref(p), cabi.ptr(hid3), cabi.ptr(mask), cabi.ptr(rec), cabi.ptr(ws), nbytes,
                       cabi.stream_ptr()), f"mvc_{cls.KIND}_recon_forward")
        ctx.dims = dims
        ctx.hid_shape = hid.shape
        ctx.save_for_backward(hid3, mask, ws, *params)
        return rec

    @classmethod
    def _backward(cls, ctx, drec):
        lib = cabi.lib()
        _, bws_b, _, bwd = cls._fns(lib, cls.KIND)
        hid3, mask, ws, *params = ctx.saved_tensors
        d, p = _rec_structs(ctx.dims, params)
        drec = _f32c(drec, "grad")
        dhid = torch.empty_like(hid3)
        grads = [_grad_like(t) for t in params]
        g = cabi.ReconGrads(*([C.c_void_p(t.data_ptr()) for t in grads] + [None] * (8 - len(grads))))
        nbytes = bws_b(C.byref(d))
        bws = cabi.workspace(nbytes, hid3.device)
        cabi.check(bwd(C.byref(d), C.byref(p), cabi.ptr(hid3), cabi.ptr(mask), cabi.ptr(drec), cabi.ptr(ws),
                       cabi.ptr(dhid), C.byref(g), cabi.ptr(bws), nbytes, cabi.stream_ptr()),
                   f"mvc_{cls.KIND}_recon_backward")
        return (None, dhid.reshape(ctx.hid_shape), None, *grads)


class GlobalReconFn(_ReconFn):
    KIND = "global"

    @staticmethod
    def forward(ctx, dims, hid, mask, *params):
        return GlobalReconFn._forward(ctx, dims, hid, mask, params)

    @staticmethod
    def backward(ctx, drec):
        return GlobalReconFn._backward(ctx, drec)


class LocalReconFn(_ReconFn):
    KIND = "local"

    @staticmethod
    def forward(ctx, dims, hid, mask, *params):
        return LocalReconFn._forward(ctx, dims, hid, mask, params)

    @staticmethod
    def backward(ctx, drec):
        return LocalReconFn._backward(ctx, drec)


# --------------------------------------------------------------------------- losses
def _slice_view(t: torch.Tensor):
    """(tensor, row pitch) for a [B,N,F] tensor whose rows are uniformly strided (a last-dim slice of a
    contiguous tensor qualifies); anything else is made contiguous."""
    if t.stride(2) == 1 and t.stride(0) == t.shape[1] * t.stride(1):
        return t, t.stride(1)
    t = t.contiguous()
    return t, t.stride(1)


class ModalityLossFn(torch.autograd.Function):
    """ModalityWiseReconstructionLoss (losses.py:86-126) forward + gradient in fused kernels.

    Returns (loss, ce, entropy, audio_rec, visual_rec); only ``loss`` carries gradient
    (the others are reported scalars, which is how train.py:195-205 uses them)."""

    @staticmethod
    def forward(ctx, output, captions, audio, arec, visual, vrec, reg_lambda, a_lambda, v_lambda, rec_type):
        lib = cabi.lib()
        L, B, V = output.shape
        x = _f32c(output.detach(), "output")
        dev = x.device
        captions = captions.to(device=dev, dtype=torch.int64).contiguous()
        need_grad = any(t is not None and t.requires_grad for t in (output, arec, vrec))
        res = torch.zeros(8, device=dev, dtype=torch.float32)
        lws = cabi.workspace(1024, dev)
        dout = torch.empty_like(x) if (need_grad and output.requires_grad) else None
        cabi.check(lib.mvc_caption_loss(cabi.ptr(x), cabi.ptr(captions), L, B, V, cabi.ptr(res), cabi.ptr(dout), 1.0,
                                        float(reg_lambda), cabi.ptr(lws), cabi.stream_ptr()), "mvc_caption_loss")
        grads = [dout, None, None]
        terms = [res[0], res[1]]
        for i, (feat, rec, lam) in enumerate(((audio, arec, a_lambda), (visual, vrec, v_lambda))):
            if rec is None or rec_type not in ("global", "local"):
                terms.append(torch.zeros((), device=dev, dtype=torch.float32))     # losses.py:100-101
                continue
            feat = _f32c(feat.detach(), "features")
            r = _f32c(rec.detach(), "features_recons") if rec.stride(2) != 1 else rec.detach()
            feat, x_ld = _slice_view(feat)
            r, r_ld = _slice_view(r)
            F = feat.shape[2]
            dr = torch.zeros(r.shape, device=dev, dtype=torch.float32) if (need_grad and rec.requires_grad) else None
            slot = res[3 + i:4 + i]
            ws_i = lws[256 * (i + 1):]
            if rec_type == "global":
                cabi.check(lib.mvc_global_recon_loss(cabi.ptr(feat), x_ld, cabi.ptr(r), r_ld, B, feat.shape[1], L, F,
                                                     cabi.ptr(captions), cabi.ptr(slot), cabi.ptr(dr), F, float(lam),
                                                     cabi.ptr(ws_i), cabi.stream_ptr()), "mvc_global_recon_loss")
            else:
                if r.shape != feat.shape:
                    raise RuntimeError(f"local reconstruction loss: shapes differ {tuple(r.shape)} vs {tuple(feat.shape)}")
                cabi.check(lib.mvc_local_recon_loss(cabi.ptr(feat), x_ld, cabi.ptr(r), r_ld, B * feat.shape[1], F,
                                                    cabi.ptr(slot), cabi.ptr(dr), F, float(lam), cabi.ptr(ws_i),
                                                    cabi.stream_ptr()), "mvc_local_recon_loss")
            terms.append(res[3 + i])
            grads[1 + i] = dr
        ce, ent, a_l, v_l = terms
        loss = ce + reg_lambda * ent + a_lambda * a_l + v_lambda * v_l                # losses.py:122-124
        ctx.save_for_backward(*[g if g is not None else torch.empty(0, device=dev) for g in grads])
        ctx.have = [g is not None for g in grads]
        ctx.mark_non_differentiable(ce, ent, a_l, v_l)
        return loss, ce, ent, a_l, v_l

    @staticmethod
    def backward(ctx, g_loss, *_unused):
        dout, da, dv = [t if h else None for t, h in zip(ctx.saved_tensors, ctx.have)]
        scale = lambda t: None if t is None else t * g_loss
        return (scale(dout), None, None, scale(da), None, scale(dv), None, None, None, None)


def clip_adam_step(param, grad, exp_avg, exp_avg_sq, max_exp_avg_sq, *, lr, betas=(0.9, 0.999), eps=1e-8,
                   weight_decay=0.0, clip_value=0.0, step=1, grad_scale=1.0):
    """clip_grad_value_ + Adam(amsgrad=True, weight_decay) (train.py:86-88, 207-210) on flat fp32 buffers."""
    lib = cabi.lib()
    cabi.check(lib.mvc_clip_adam_step(cabi.ptr(param), cabi.ptr(grad), cabi.ptr(exp_avg), cabi.ptr(exp_avg_sq),
                                      cabi.ptr(max_exp_avg_sq), param.numel(), lr, betas[0], betas[1], eps,
                                      weight_decay, clip_value, int(step), grad_scale, cabi.stream_ptr()),
               "mvc_clip_adam_step")
