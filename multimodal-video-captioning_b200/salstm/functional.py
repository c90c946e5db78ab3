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


# Gradient arena (salstm/trainer.py::FlatClipAdam): the optimizer keeps all gradients in one flat buffer and
# registers (parameter -> view of that buffer) here.  A backward kernel may write a parameter's gradient straight
# into its view -- autograd then adopts the view as .grad: no zero-fill, no accumulate pass -- but ONLY when that is
# indistinguishable from returning a fresh tensor: the parameter's .grad is None (otherwise autograd would do
# `p.grad += view` on a tensor aliasing itself) and the view has not been handed out before in this step (a
# parameter feeding two nodes).  Everything else (gradient accumulation, zero_grad(set_to_none=False), a second
# backward with retain_graph) gets a fresh tensor and autograd accumulates as usual.
import weakref


class GradArena:
    def __init__(self, params, views):
        self._by_ptr = {p.data_ptr(): (weakref.ref(p), v) for p, v in zip(params, views)}
        self._handed = set()

    def new_step(self):
        self._handed.clear()

    def take(self, t: torch.Tensor):
        ent = self._by_ptr.get(t.data_ptr())
        if ent is None:
            return None
        p, v = ent[0](), ent[1]
        if p is None or p.grad is not None or t.data_ptr() in self._handed or v.shape != t.shape or v.device != t.device:
            return None
        self._handed.add(t.data_ptr())
        return v.detach()          # a fresh alias: autograd may adopt it as .grad without cloning


_GRAD_ARENAS: "weakref.WeakSet[GradArena]" = weakref.WeakSet()


def register_grad_arena(arena: GradArena) -> None:
    """Scoped per optimizer: the arena lives as long as the optimizer that registered it."""
    _GRAD_ARENAS.add(arena)


def _grad_like(t: torch.Tensor) -> torch.Tensor:
    for arena in _GRAD_ARENAS:
        v = arena.take(t)
        if v is not None:
            return v
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
                                               cabi.ptr(ws), nbytes, int(any(ctx.needs_input_grad)), cabi.stream_ptr()),
                       "mvc_decoder_forward")      # (a backward pass will follow: forward also prepares its operands)
        ctx.dims = dims
        ctx.save_for_backward(out, tokens, ws, *params)
        # an output nothing consumed (the hidden states without a reconstructor) arrives as None in backward instead of a
        # zero tensor that would be filled and added for nothing
        ctx.set_materialize_grads(False)
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


# --------------------------------------------------------------------------- single-step blocks (differentiable)
# forward_word (features_captioning.py:77-89) and the stand-alone TemporalAttention.forward
# (temporal_attention.py:19-33) outside the fused time loop: the same block kernels the fp32 launch
# chain uses, each wrapped as an autograd node so that a caller stepping the decoder word by word
# gets the reference's gradients.
def _gemm(M, N, K, a, a_rs, a_cs, b, b_rs, b_cs, c, ldc, beta=0.0, bias=None, c_off=0):
    """c[M,N] (ld ldc, element offset c_off) = beta*c + bias + sum_k a(m,k) b(n,k); fp32 FFMA kernel."""
    cp = C.c_void_p(c.data_ptr() + 4 * c_off)
    cabi.check(cabi.lib().mvc_gemm_f32(M, N, K, 1.0, cabi.ptr(a), a_rs, a_cs, cabi.ptr(b), b_rs, b_cs, beta, cp, ldc,
                                       cabi.ptr(bias), cabi.stream_ptr()), "mvc_gemm_f32")


class SoftAttentionFn(torch.autograd.Function):
    """(ctx [B,F], alpha [B,T]) = TemporalAttention(hidden, feats, masks); temporal_attention.py:19-33.
    alpha is returned for inspection only (non-differentiable, as every caller of the reference uses it)."""

    @staticmethod
    def forward(ctx, hidden, feats, masks, W, U, b, w):
        lib = cabi.lib()
        hidden, feats = _f32c(hidden.detach(), "hidden"), _f32c(feats.detach(), "feats")
        W, U, b, w = (_f32c(t.detach(), "parameter") for t in (W, U, b, w))
        B, T, F = feats.shape
        A, H = W.shape
        dev, st = feats.device, cabi.stream_ptr()
        wq = torch.empty(B, A, device=dev)
        uk = torch.empty(B * T, A, device=dev)
        _gemm(B, A, H, hidden, H, 1, W, H, 1, wq, A)
        _gemm(B * T, A, F, feats, F, 1, U, F, 1, uk, A)
        out = torch.empty(B, F, device=dev)
        alpha = torch.empty(B, T, device=dev)
        m = None if masks is None else masks.to(device=dev, dtype=torch.uint8).contiguous()
        cabi.check(lib.mvc_soft_attention_fwd(B, T, A, F, cabi.ptr(wq), cabi.ptr(uk), cabi.ptr(b), cabi.ptr(w),
                                              cabi.ptr(feats), 0, B, T * F, F, cabi.ptr(m), T, 1, cabi.ptr(out), F, None, 0,
                                              cabi.ptr(alpha), 0, st), "mvc_soft_attention_fwd")
        ctx.save_for_backward(hidden, feats, W, U, b, w, wq, uk, alpha)
        ctx.mark_non_differentiable(alpha)
        return out, alpha

    @staticmethod
    def backward(ctx, dctx, _dalpha):
        lib = cabi.lib()
        hidden, feats, W, U, b, w, wq, uk, alpha = ctx.saved_tensors
        B, T, F = feats.shape
        A, H = W.shape
        dev, st = feats.device, cabi.stream_ptr()
        dctx = _f32c(dctx, "grad")
        dwq = torch.empty(B, A, device=dev)
        duk = torch.zeros(B * T, A, device=dev)
        dwp = torch.zeros(B, A, device=dev)
        dfeats = torch.zeros(B, T, F, device=dev)
        cabi.check(lib.mvc_soft_attention_bwd(B, T, A, F, cabi.ptr(wq), cabi.ptr(uk), cabi.ptr(b), cabi.ptr(w),
                                              cabi.ptr(feats), 0, T * F, F, cabi.ptr(alpha), cabi.ptr(dctx), F,
                                              cabi.ptr(dwq), cabi.ptr(duk), cabi.ptr(dwp), cabi.ptr(dfeats), T * F, F, 0,
                                              st), "mvc_soft_attention_bwd")
        dhidden = torch.empty(B, H, device=dev)
        dW, dU = torch.empty_like(W), torch.empty_like(U)
        db, dw = torch.empty_like(b), torch.empty_like(w)
        _gemm(B, H, A, dwq, A, 1, W, 1, H, dhidden, H)                     # dwq . W
        _gemm(A, H, B, dwq, 1, A, hidden, 1, H, dW, H)                      # dwq^T . hidden
        _gemm(A, F, B * T, duk, 1, A, feats, 1, F, dU, F)                   # duk^T . feats
        _gemm(B * T, F, A, duk, A, 1, U, 1, F, dfeats, F, beta=1.0)         # dfeats += duk . U
        cabi.check(lib.mvc_colsum(cabi.ptr(dwq), B, A, A, cabi.ptr(db), st), "mvc_colsum")
        cabi.check(lib.mvc_colsum(cabi.ptr(dwp), B, A, A, cabi.ptr(dw), st), "mvc_colsum")
        return dhidden, dfeats, None, dW, dU, db, dw


class WordStepFn(torch.autograd.Function):
    """One decoder step after the attention: embedding -> nn.LSTM step on [emb ; ctx] -> out -> log_softmax
    (features_captioning.py:78-88) -> (log_probs [B,V], h1 [B,H], c1 [B,H])."""

    @staticmethod
    def forward(ctx, words, ctxv, h0, c0, emb_w, w_ih, w_hh, b_ih, b_hh, out_w, out_b):
        lib = cabi.lib()
        st = cabi.stream_ptr()
        ctxv, h0, c0 = (_f32c(t.detach(), "state") for t in (ctxv, h0, c0))
        emb_w, w_ih, w_hh, b_ih, b_hh, out_w, out_b = (_f32c(t.detach(), "parameter")
                                                       for t in (emb_w, w_ih, w_hh, b_ih, b_hh, out_w, out_b))
        B, F = ctxv.shape
        V, E = emb_w.shape
        H = w_hh.shape[1]
        dev = ctxv.device
        words = words.reshape(-1).to(device=dev, dtype=torch.int64).contiguous()
        emb = torch.empty(B, E, device=dev)
        cabi.check(lib.mvc_embedding_gather(cabi.ptr(emb_w), E, cabi.ptr(words), B, cabi.ptr(emb), E, 0, st), "gather")
        pre = torch.empty(B, 4 * H, device=dev)
        w_ctx = C.c_void_p(w_ih.data_ptr() + 4 * E)
        cabi.check(lib.mvc_gemm_f32(B, 4 * H, E, 1.0, cabi.ptr(emb), E, 1, cabi.ptr(w_ih), E + F, 1, 0.0, cabi.ptr(pre),
                                    4 * H, cabi.ptr(b_ih), st), "mvc_gemm_f32")
        cabi.check(lib.mvc_gemm_f32(B, 4 * H, F, 1.0, cabi.ptr(ctxv), F, 1, w_ctx, E + F, 1, 1.0, cabi.ptr(pre), 4 * H,
                                    cabi.ptr(b_hh), st), "mvc_gemm_f32")
        _gemm(B, 4 * H, H, h0, H, 1, w_hh, H, 1, pre, 4 * H, beta=1.0)
        act = torch.empty(B, 4 * H, device=dev)
        h1, c1 = torch.empty(B, H, device=dev), torch.empty(B, H, device=dev)
        cabi.check(lib.mvc_lstm_cell_fwd(B, H, cabi.ptr(pre), None, 0, None, None, None, cabi.ptr(c0), cabi.ptr(act),
                                         cabi.ptr(c1), cabi.ptr(h1), H, None, 0, None, 0, st), "mvc_lstm_cell_fwd")
        logp = torch.empty(B, V, device=dev)
        _gemm(B, V, H, h1, H, 1, out_w, H, 1, logp, V, bias=out_b)
        cabi.check(lib.mvc_log_softmax_rows(cabi.ptr(logp), B, V, None, st), "mvc_log_softmax_rows")
        ctx.save_for_backward(words, emb, ctxv, h0, c0, act, c1, h1, logp, emb_w, w_ih, w_hh, out_w)
        return logp, h1, c1

    @staticmethod
    def backward(ctx, dlogp, dh1, dc1):
        lib = cabi.lib()
        st = cabi.stream_ptr()
        words, emb, ctxv, h0, c0, act, c1, h1, logp, emb_w, w_ih, w_hh, out_w = ctx.saved_tensors
        B, F = ctxv.shape
        V, E = emb_w.shape
        H = w_hh.shape[1]
        dev = ctxv.device
        dh = torch.zeros(B, H, device=dev) if dh1 is None else _f32c(dh1, "grad").clone()
        dout_w, dout_b = torch.zeros_like(out_w), torch.zeros(V, device=dev)
        if dlogp is not None:
            dlogits = torch.empty(B, V, device=dev)
            cabi.check(lib.mvc_log_softmax_bwd(cabi.ptr(logp), cabi.ptr(_f32c(dlogp, "grad")), B, V, cabi.ptr(dlogits),
                                               None, st), "mvc_log_softmax_bwd")
            _gemm(B, H, V, dlogits, V, 1, out_w, 1, H, dh, H, beta=1.0)     # dh += dlogits . out_w
            _gemm(V, H, B, dlogits, 1, V, h1, 1, H, dout_w, H)              # dlogits^T . h1
            cabi.check(lib.mvc_colsum(cabi.ptr(dlogits), B, V, V, cabi.ptr(dout_b), st), "mvc_colsum")
        dc = torch.zeros(B, H, device=dev) if dc1 is None else _f32c(dc1, "grad").clone()
        dg = torch.empty(B, 4 * H, device=dev)
        cabi.check(lib.mvc_lstm_cell_bwd(B, H, cabi.ptr(act), cabi.ptr(c0), cabi.ptr(c1), cabi.ptr(dh), H, None, 0,
                                         cabi.ptr(dc), cabi.ptr(dg), None, st), "mvc_lstm_cell_bwd")
        K = E + F
        demb, dctx, dh0 = torch.empty(B, E, device=dev), torch.empty(B, F, device=dev), torch.empty(B, H, device=dev)
        _gemm(B, E, 4 * H, dg, 4 * H, 1, w_ih, 1, K, demb, E)
        cabi.check(lib.mvc_gemm_f32(B, F, 4 * H, 1.0, cabi.ptr(dg), 4 * H, 1, C.c_void_p(w_ih.data_ptr() + 4 * E), 1, K,
                                    0.0, cabi.ptr(dctx), F, None, st), "mvc_gemm_f32")
        _gemm(B, H, 4 * H, dg, 4 * H, 1, w_hh, 1, H, dh0, H)
        dw_ih, dw_hh = torch.empty_like(w_ih), torch.empty_like(w_hh)
        _gemm(4 * H, E, B, dg, 1, 4 * H, emb, 1, E, dw_ih, K)
        _gemm(4 * H, F, B, dg, 1, 4 * H, ctxv, 1, F, dw_ih, K, c_off=E)
        _gemm(4 * H, H, B, dg, 1, 4 * H, h0, 1, H, dw_hh, H)
        db = torch.empty(4 * H, device=dev)
        cabi.check(lib.mvc_colsum(cabi.ptr(dg), B, 4 * H, 4 * H, cabi.ptr(db), st), "mvc_colsum")
        demb_w = torch.zeros_like(emb_w)
        cabi.check(lib.mvc_embedding_scatter_add(cabi.ptr(demb), E, E, cabi.ptr(words), B, cabi.ptr(demb_w), st), "scatter")
        return None, dctx, dh0, dc, demb_w, dw_ih, dw_hh, db, db.clone(), dout_w, dout_b


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


class SplitFeaturesFn(torch.autograd.Function):
    """rec [B,N,Fa+Fv] -> (rec[..., :Fa], rec[..., Fa:]) (captioning.py:125-126) as views; the backward joins the two
    gradients with one concat kernel instead of autograd's two zero-fills + two copies + add."""

    @staticmethod
    def forward(ctx, rec, Fa):
        ctx.Fa, ctx.shape = Fa, rec.shape
        return rec[:, :, :Fa], rec[:, :, Fa:]

    @staticmethod
    def backward(ctx, da, dv):
        B, N, F = ctx.shape
        Fa = ctx.Fa
        dev = (da if da is not None else dv).device
        da = torch.zeros(B, N, Fa, device=dev) if da is None else _f32c(da, "grad")
        dv = torch.zeros(B, N, F - Fa, device=dev) if dv is None else _f32c(dv, "grad")
        out = torch.empty(B, N, F, device=dev, dtype=torch.float32)
        cabi.check(cabi.lib().mvc_concat_cast(cabi.ptr(da), Fa, cabi.ptr(dv), F - Fa, B * N, cabi.ptr(out), 0,
                                              cabi.stream_ptr()), "mvc_concat_cast")
        return out, None


# --------------------------------------------------------------------------- losses
def _slice_view(t: torch.Tensor):
    """(tensor, row pitch) for a [B,N,F] tensor whose rows are uniformly strided (a last-dim slice of a
    contiguous tensor qualifies); anything else is made contiguous."""
    if t.stride(2) == 1 and t.stride(0) == t.shape[1] * t.stride(1):
        return t, t.stride(1)
    t = t.contiguous()
    return t, t.stride(1)


class ModalityLossFn(torch.autograd.Function):
    """ModalityWiseReconstructionLoss (losses.py:86-126): value and gradient in fused kernels.

    Returns (loss, ce, entropy, audio_rec, visual_rec) as views of ONE device buffer (no torch arithmetic).
    The gradient of `loss` w.r.t. the log-probs / reconstructions is produced by the forward pass and only
    scaled by the upstream gradient in backward (a kernel that exits at once when that gradient is 1, as in
    train.py:198).  Backward through the individual terms (`ce.backward()`, ...) is supported like the
    reference's graph tensors, by re-running the loss kernels with the matching scales (rare path, one host
    read of the four upstream scalars)."""

    @staticmethod
    def forward(ctx, output, captions, audio, arec, visual, vrec, reg_lambda, a_lambda, v_lambda, rec_type):
        x = _f32c(output.detach(), "output")
        captions = captions.to(device=x.device, dtype=torch.int64).contiguous()
        recs = []
        for feat, rec in ((audio, arec), (visual, vrec)):
            if rec is None or rec_type not in ("global", "local"):
                recs.append(None)                                                    # losses.py:100-101
                continue
            feat = _f32c(feat.detach(), "features")
            r = _f32c(rec.detach(), "features_recons") if rec.stride(2) != 1 else rec.detach()
            if rec_type == "local" and r.shape != feat.shape:
                raise RuntimeError(f"local reconstruction loss: shapes differ {tuple(r.shape)} vs {tuple(feat.shape)}")
            recs.append((_slice_view(feat), _slice_view(r)))
        want = [output.requires_grad, arec is not None and arec.requires_grad and recs[0] is not None,
                vrec is not None and vrec.requires_grad and recs[1] is not None]
        res, grads = ModalityLossFn._run(x, captions, recs, rec_type, want, 1.0, float(reg_lambda),
                                         (float(a_lambda), float(v_lambda)), (reg_lambda, a_lambda, v_lambda))
        ctx.set_materialize_grads(False)
        ctx.cfg = (rec_type, want, float(reg_lambda), float(a_lambda), float(v_lambda))
        ctx.recs = recs
        ctx.scaled = False
        ctx.save_for_backward(x, captions, *[g if g is not None else torch.empty(0, device=x.device) for g in grads])
        return res[5], res[0], res[1], res[3], res[4]

    @staticmethod
    def _run(x, captions, recs, rec_type, want, ce_scale, ent_scale, rec_scales, lambdas):
        """Launch the loss kernels: -> (res [8] fp32: ce, entropy, n_tokens, a_rec, v_rec, total; gradients)."""
        lib = cabi.lib()
        st = cabi.stream_ptr()
        L, B, V = x.shape
        dev = x.device
        res = torch.empty(8, device=dev, dtype=torch.float32)
        lws = cabi.workspace(1024, dev)
        dout = torch.empty_like(x) if want[0] else None
        cabi.check(lib.mvc_caption_loss(cabi.ptr(x), cabi.ptr(captions), L, B, V, cabi.ptr(res), cabi.ptr(dout),
                                        float(ce_scale), float(ent_scale), cabi.ptr(lws), st), "mvc_caption_loss")
        grads = [dout, None, None]
        for i, pair in enumerate(recs):
            if pair is None:
                continue
            (feat, x_ld), (r, r_ld) = pair
            F = feat.shape[2]
            dr = torch.empty(r.shape, device=dev, dtype=torch.float32) if want[1 + i] else None
            slot = res[3 + i:4 + i]
            ws_i = lws[256 * (i + 1):]
            if rec_type == "global":
                cabi.check(lib.mvc_global_recon_loss(cabi.ptr(feat), x_ld, cabi.ptr(r), r_ld, B, feat.shape[1], L, F,
                                                     cabi.ptr(captions), cabi.ptr(slot), cabi.ptr(dr), F,
                                                     float(rec_scales[i]), cabi.ptr(ws_i), st), "mvc_global_recon_loss")
            else:
                cabi.check(lib.mvc_local_recon_loss(cabi.ptr(feat), x_ld, cabi.ptr(r), r_ld, B * feat.shape[1], F,
                                                    cabi.ptr(slot), cabi.ptr(dr), F, float(rec_scales[i]), cabi.ptr(ws_i),
                                                    st), "mvc_local_recon_loss")
            grads[1 + i] = dr
        cabi.check(lib.mvc_loss_combine(cabi.ptr(res), float(lambdas[0]), float(lambdas[1]), float(lambdas[2]),
                                        int(recs[0] is not None), int(recs[1] is not None), st), "mvc_loss_combine")
        return res, grads

    @staticmethod
    def backward(ctx, g_loss, g_ce, g_ent, g_a, g_v):
        lib = cabi.lib()
        x, captions, *saved = ctx.saved_tensors
        rec_type, want, reg, a_l, v_l = ctx.cfg
        grads = [t if w else None for t, w in zip(saved, want)]
        if g_ce is None and g_ent is None and g_a is None and g_v is None:
            if g_loss is None:
                return (None,) * 10
            if ctx.scaled:
                raise RuntimeError("mvc_b200: second backward through the fused loss (its gradient buffers were "
                                   "scaled in place by the first); recompute the loss instead of retain_graph")
            ctx.scaled = True
            g = _f32c(g_loss, "grad").reshape(1)
            for t in grads:
                if t is not None:
                    cabi.check(lib.mvc_scale_by_scalar(cabi.ptr(t), t.numel(), cabi.ptr(g), cabi.stream_ptr()),
                               "mvc_scale_by_scalar")
        else:
            # gradient requested through ce / entropy / reconstruction terms themselves: fold the upstream scalars
            # into the kernel scales and run the kernels again
            gl, gc, ge, ga, gv = (0.0 if t is None else float(t) for t in (g_loss, g_ce, g_ent, g_a, g_v))
            _, grads = ModalityLossFn._run(x, captions, ctx.recs, rec_type, want, gl + gc, gl * reg + ge,
                                           (gl * a_l + ga, gl * v_l + gv), (reg, a_l, v_l))
        return (grads[0], None, None, grads[1], None, grads[2], None, None, None, None)


def clip_adam_step(param, grad, exp_avg, exp_avg_sq, max_exp_avg_sq, *, lr, betas=(0.9, 0.999), eps=1e-8,
                   weight_decay=0.0, clip_value=0.0, step=1, grad_scale=1.0):
    """clip_grad_value_ + Adam(amsgrad=True, weight_decay) (train.py:86-88, 207-210) on flat fp32 buffers."""
    lib = cabi.lib()
    cabi.check(lib.mvc_clip_adam_step(cabi.ptr(param), cabi.ptr(grad), cabi.ptr(exp_avg), cabi.ptr(exp_avg_sq),
                                      cabi.ptr(max_exp_avg_sq), param.numel(), lr, betas[0], betas[1], eps,
                                      weight_decay, clip_value, int(step), grad_scale, cabi.stream_ptr()),
               "mvc_clip_adam_step")
