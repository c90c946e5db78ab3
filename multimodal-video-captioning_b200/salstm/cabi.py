"""ctypes binding of libmvc_b200.so (include/mvc_b200.h).

The library is the product: there is no Python/ATen fallback.  If the shared
object is missing, or the process has no CUDA device, every compute entry
raises ``RuntimeError`` -- loudly, never silently.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "lib", "libmvc_b200.so")

F32, BF16 = 0, 1
MVC_F32, MVC_BF16 = F32, BF16
MVC_INPUT_F32, MVC_INPUT_BF16 = 0, 1      # mvc_set_input_format (include/mvc_b200.h)
PRECISIONS = {"fp32": F32, "f32": F32, "float32": F32, "bf16": BF16, "bfloat16": BF16}

vp, i32, i64, f32, sz = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_size_t
P = C.POINTER


class DecoderDims(C.Structure):
    _fields_ = [(n, i32) for n in ("B", "T", "F", "H", "E", "A", "V", "L", "precision")]


_DEC_PARAM_FIELDS = ("embedding", "att_W", "att_U", "att_b", "att_w", "w_ih", "w_hh", "b_ih", "b_hh", "out_w", "out_b")


class DecoderParams(C.Structure):
    _fields_ = [(n, vp) for n in _DEC_PARAM_FIELDS]


class DecoderGrads(C.Structure):
    _fields_ = [(n, vp) for n in _DEC_PARAM_FIELDS]


class ReconDims(C.Structure):
    _fields_ = [(n, i32) for n in ("B", "L", "H", "Fr", "A", "T", "precision")]


_REC_PARAM_FIELDS = ("w_ih", "w_hh", "b_ih", "b_hh", "att_W", "att_U", "att_b", "att_w")


class ReconParams(C.Structure):
    _fields_ = [(n, vp) for n in _REC_PARAM_FIELDS]


class ReconGrads(C.Structure):
    _fields_ = [(n, vp) for n in _REC_PARAM_FIELDS]


# name -> (restype, argtypes).  Every symbol declared in include/mvc_b200.h appears here;
# tests/test_cabi_symbols.py cross-checks the two lists.
SIGNATURES = {
    "mvc_last_error": (C.c_char_p, []),
    "mvc_version": (i32, []),
    "mvc_device_ok": (i32, []),
    "mvc_launch_count": (C.c_longlong, []),
    "mvc_prof_arm": (i32, [i32, i32, i32, i32]),
    "mvc_prof_collect": (i32, [P(C.c_double), P(C.c_longlong)]),
    "mvc_debug_set_recur_prof": (i32, [vp]),
    "mvc_debug_set_recur_bwd_prof": (i32, [vp]),
    "mvc_debug_set_attn_prof": (i32, [vp]),
    "mvc_debug_set_gemm_prof": (i32, [vp, i32, i32, i32]),
    "mvc_gemm_f32": (i32, [i32, i32, i32, f32, vp, i64, i64, vp, i64, i64, f32, vp, i64, vp, vp]),
    "mvc_gemm_bf16": (i32, [i32, i32, i32, vp, i64, vp, i64, f32, vp, i64, vp, vp, i64, vp]),
    "mvc_gemm_bf16_ex": (i32, [i32, i32, i32, vp, i64, i32, vp, i64, i32, f32, vp, i64, vp, vp]),
    "mvc_concat_cast": (i32, [vp, i32, vp, i32, i64, vp, i32, vp]),
    "mvc_set_input_format": (i32, [i32]),
    "mvc_get_input_format": (i32, []),
    "mvc_concat_bf16": (i32, [vp, i32, vp, i32, i64, vp, vp]),
    "mvc_cast_bf16": (i32, [vp, vp, i64, vp]),
    "mvc_transpose_to_bf16": (i32, [vp, i32, i64, i64, i64, vp, i64, vp]),
    "mvc_soft_attention_fwd": (i32, [i32, i32, i32, i32, vp, vp, vp, vp, vp, i32, i32, i64, i64, vp, i64, i64, vp, i64,
                                     vp, i64, vp, i32, vp]),
    "mvc_soft_attention_bwd": (i32, [i32, i32, i32, i32, vp, vp, vp, vp, vp, i32, i64, i64, vp, vp, i64, vp, vp, vp, vp,
                                     i64, i64, i32, vp]),
    "mvc_lstm_cell_fwd": (i32, [i32, i32, vp, vp, i64, vp, vp, vp, vp, vp, vp, vp, i64, vp, i64, vp, i64, vp]),
    "mvc_pack_gate_rows_bf16": (i32, [vp, i32, i32, i64, i32, vp, vp]),
    "mvc_lstm_gates_cell_bf16": (i32, [i32, i32, i32, vp, i64, vp, i64, vp, vp, i64, vp, vp, vp, vp, i64, vp, i64, vp]),
    "mvc_vocab_argmax_bf16": (i32, [i32, i32, i32, vp, i64, vp, i64, vp, vp, sz, vp, vp]),
    "mvc_vocab_aux_row0": (i32, [i32]),
    "mvc_vocab_argmax_wq_bf16": (i32, [i32, i32, i32, i32, vp, i64, vp, i64, vp, vp, sz, vp, vp, vp]),
    "mvc_vocab_topk_workspace_bytes": (sz, [i32, i32]),
    "mvc_vocab_topk_bf16": (i32, [i32, i32, i32, vp, i64, vp, i64, vp, i32, vp, sz, vp, vp, vp]),
    "mvc_lstm_cell_bwd": (i32, [i32, i32, vp, vp, vp, vp, i64, vp, i64, vp, vp, vp, vp]),
    "mvc_log_softmax_rows": (i32, [vp, i64, i32, vp, vp]),
    "mvc_argmax_rows": (i32, [vp, vp, i64, i32, vp, vp]),
    "mvc_log_softmax_bwd": (i32, [vp, vp, i64, i32, vp, vp, vp]),
    "mvc_embedding_gather": (i32, [vp, i32, vp, i64, vp, i64, i32, vp]),
    "mvc_embedding_scatter_add": (i32, [vp, i64, i32, vp, i64, vp, vp]),
    "mvc_colsum": (i32, [vp, i64, i32, i64, vp, vp]),
    "mvc_decoder_fwd_workspace_bytes": (sz, [P(DecoderDims), i32]),
    "mvc_decoder_bwd_workspace_bytes": (sz, [P(DecoderDims)]),
    "mvc_decoder_forward": (i32, [P(DecoderDims), P(DecoderParams), vp, i32, vp, i32, vp, vp, vp, vp, vp, vp, sz, i32, vp]),
    "mvc_decoder_backward": (i32, [P(DecoderDims), P(DecoderParams), vp, vp, vp, vp, vp, P(DecoderGrads), vp, sz, vp]),
    "mvc_decoder_greedy_workspace_bytes": (sz, [P(DecoderDims)]),
    "mvc_decoder_greedy": (i32, [P(DecoderDims), P(DecoderParams), vp, i32, vp, i32, vp, vp, sz, vp]),
    "mvc_decoder_beam_workspace_bytes": (sz, [P(DecoderDims), i32]),
    "mvc_decoder_beam": (i32, [P(DecoderDims), P(DecoderParams), vp, i32, vp, i32, i32, f32, vp, vp, sz, vp]),
    "mvc_caption_mask": (i32, [vp, i64, vp, vp]),
    "mvc_global_recon_workspace_bytes": (sz, [P(ReconDims)]),
    "mvc_global_recon_bwd_workspace_bytes": (sz, [P(ReconDims)]),
    "mvc_global_recon_forward": (i32, [P(ReconDims), P(ReconParams), vp, vp, vp, vp, sz, vp]),
    "mvc_global_recon_backward": (i32, [P(ReconDims), P(ReconParams), vp, vp, vp, vp, vp, P(ReconGrads), vp, sz, vp]),
    "mvc_local_recon_workspace_bytes": (sz, [P(ReconDims)]),
    "mvc_local_recon_bwd_workspace_bytes": (sz, [P(ReconDims)]),
    "mvc_local_recon_forward": (i32, [P(ReconDims), P(ReconParams), vp, vp, vp, vp, sz, vp]),
    "mvc_local_recon_backward": (i32, [P(ReconDims), P(ReconParams), vp, vp, vp, vp, vp, P(ReconGrads), vp, sz, vp]),
    "mvc_caption_loss_workspace_bytes": (sz, [i32, i32, i32]),
    "mvc_caption_loss": (i32, [vp, vp, i32, i32, i32, vp, vp, f32, f32, vp, vp]),
    "mvc_global_recon_loss_workspace_bytes": (sz, [i32, i32]),
    "mvc_global_recon_loss": (i32, [vp, i64, vp, i64, i32, i32, i32, i32, vp, vp, vp, i64, f32, vp, vp]),
    "mvc_local_recon_loss": (i32, [vp, i64, vp, i64, i64, i32, vp, vp, i64, f32, vp, vp]),
    "mvc_loss_combine": (i32, [vp, f32, f32, f32, i32, i32, vp]),
    "mvc_scale_by_scalar": (i32, [vp, i64, vp, vp]),
    "mvc_clip_adam_multimem": (i32, [vp, vp, vp, vp, vp, vp, i64, i64, vp, i32, f32, f32, f32, f32, f32, f32, vp]),
    "mvc_clip_adam_p2p_multimem": (i32, [vp, vp, vp, i32, vp, vp, vp, i64, i64, vp, i32, f32, f32, f32, f32, f32, f32, vp]),
    "mvc_clip_adam_step_dev": (i32, [vp, vp, vp, vp, vp, i64, vp, i32, f32, f32, f32, f32, f32, f32, vp]),
    "mvc_clip_adam_step": (i32, [vp, vp, vp, vp, vp, i64, f32, f32, f32, f32, f32, f32, i32, f32, vp]),
}

_lib = None
_lock = threading.Lock()


def load():
    """dlopen libmvc_b200.so and install the signatures (no GPU needed)."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise RuntimeError(
                    f"libmvc_b200.so not found at {LIB_PATH}: build it with "
                    "`python -c 'import __graft_entry__ as g; g.build()'` (there is no CPU/ATen fallback)")
            lib = C.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(lib, name)
                fn.restype, fn.argtypes = res, args
            _lib = lib
    return _lib


def lib():
    """The loaded library, for compute: also insists on a CUDA device."""
    if not torch.cuda.is_available():
        raise RuntimeError("mvc_b200: no CUDA device visible; this path has no CPU fallback")
    return load()


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().mvc_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"libmvc_b200 {what} failed: {msg}")


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def precision_id(p) -> int:
    if isinstance(p, int):
        return p
    try:
        return PRECISIONS[str(p).lower()]
    except KeyError:
        raise ValueError(f"unknown precision {p!r}; expected one of {sorted(PRECISIONS)}")


def workspace(nbytes: int, device) -> torch.Tensor:
    return torch.empty(int(nbytes), dtype=torch.uint8, device=device)
