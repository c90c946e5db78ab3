"""CPU oracle for the SA-LSTM caption hot path.  TEST INFRASTRUCTURE ONLY.

This file restates, as pure functions over a flat ``{name: tensor}`` parameter
dict, the arithmetic of the reference's hot path (hmartelb/multimodal-video-
captioning).  It is *not* part of the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it, and only as the checker / the CPU baseline.
The product path (``multimodal-video-captioning_b200/``) never imports it and
fails loudly when the CUDA extension is missing.

Parity pinning: the reference ships no tests or golden vectors (SURVEY.md §4,
§8c), so this oracle is pinned against outputs of the reference modules
themselves, run in the build container by ``tools/make_golden.py`` and
committed under ``tests/golden/*.npz``; ``tests/test_oracle_golden.py``
checks every function below against them.

The arithmetic lives in a third-party dependency (PyTorch; the reference pins
``torch==1.12.1`` at requirements.txt:9, this image has 2.11).  The semantics
relied on -- ``nn.LSTM`` gate order i,f,g,o with two biases, ``nll_loss`` mean
over non-ignored targets, ``mse_loss`` mean over all elements -- are restated
in closed form here (``lstm_cell``) and optionally routed through the same
ATen op the reference calls (``aten_lstm=True`` -> ``torch._VF.lstm`` on a
length-1 sequence) so the CPU baseline pays the same library cost as the
reference does.

All functions are dtype-agnostic: run them in float64 for a "ground truth"
that both the reference fp32 CPU result and the CUDA fp32 result can be
measured against.

Reference citations are ``file:line`` relative to ``/root/reference/``.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch

Params = Dict[str, torch.Tensor]

PAD, SOS, EOS = 0, 1, 2  # src/get_loader.py:25-26


# --------------------------------------------------------------------------
# building blocks
# --------------------------------------------------------------------------
def lstm_cell(x, h, c, w_ih, w_hh, b_ih, b_hh, aten_lstm: bool = False):
    """One LSTM step.  Reference: ``self.rnn(input_combined, hidden)``
    src/models/features_captioning.py:84 (nn.LSTM, 1 layer, 1 direction).

    gates = x W_ih^T + b_ih + h W_hh^T + b_hh, split in PyTorch order
    (i, f, g, o); c' = sig(f) c + sig(i) tanh(g); h' = sig(o) tanh(c').
    """
    if aten_lstm:
        _, hn, cn = torch._VF.lstm(
            x.unsqueeze(0), (h.unsqueeze(0), c.unsqueeze(0)),
            [w_ih, w_hh, b_ih, b_hh], True, 1, 0.0, False, False, False)
        return hn[0], cn[0]
    gates = x @ w_ih.t() + b_ih + h @ w_hh.t() + b_hh
    i, f, g, o = gates.chunk(4, dim=1)
    c2 = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
    h2 = torch.sigmoid(o) * torch.tanh(c2)
    return h2, c2


def soft_attention(p: Params, pre: str, query, keys, mask=None, keys_proj=None):
    """Additive attention.  Reference: TemporalAttention.forward
    src/models/temporal_attention.py:19-33.

    e[b,t] = w . tanh(W q[b] + U k[b,t] + bias);  e[~mask] = -inf (optional);
    alpha = softmax_t(e);  ctx[b] = sum_t alpha[b,t] k[b,t].
    ``keys_proj`` lets a caller hoist ``U k`` (loop invariant); when None it is
    recomputed here exactly like the reference does every timestep (:21).
    Returns (ctx [B,F], alpha [B,T]).
    """
    wq = query @ p[pre + "W.weight"].t()                                 # :20
    uk = keys @ p[pre + "U.weight"].t() if keys_proj is None else keys_proj  # :21
    e = torch.tanh(wq.unsqueeze(1) + uk + p[pre + "b"]) @ p[pre + "w.weight"][0]  # :22-23
    if mask is not None:
        e = e.masked_fill(~mask, -float("inf"))                          # :25-28
    alpha = torch.softmax(e, dim=1)                                      # :29
    ctx = (keys * alpha.unsqueeze(2)).sum(dim=1)                         # :31-32
    return ctx, alpha


# --------------------------------------------------------------------------
# decoder (FeaturesCaptioning)
# --------------------------------------------------------------------------
def decoder_step(p: Params, pre: str, feats, h, c, prev_words, *, hoisted_uv=None,
                 aten_lstm=False):
    """Reference: FeaturesCaptioning.forward_word
    src/models/features_captioning.py:77-89.  LSTM input is [embedding ; ctx]
    (:83).  Returns (log_probs [B,V], h', c', alpha [B,T])."""
    emb = p[pre + "embedding.weight"][prev_words]                        # :78
    ctx, alpha = soft_attention(p, pre + "attention.", h, feats, keys_proj=hoisted_uv)  # :80-81
    x = torch.cat([emb, ctx], dim=1)                                     # :83
    h2, c2 = lstm_cell(x, h, c, p[pre + "rnn.weight_ih_l0"], p[pre + "rnn.weight_hh_l0"],
                       p[pre + "rnn.bias_ih_l0"], p[pre + "rnn.bias_hh_l0"], aten_lstm)
    logits = h2 @ p[pre + "out.weight"].t() + p[pre + "out.bias"]        # :87
    return torch.log_softmax(logits, dim=1), h2, c2, alpha               # :88


def teacher_flags(captions, max_len: int, tf_ratio: float) -> List[bool]:
    """Per-step teacher-forcing decisions, drawn from the global CPU RNG in
    the same order and number as the reference: one ``torch.rand(1)`` per
    loop step iff captions are given (features_captioning.py:113-116)."""
    if captions is None:
        return [False] * (max_len - 1)
    return [bool(torch.rand(1) < tf_ratio) for _ in range(1, max_len)]


def decoder_decode(p: Params, pre: str, feats, captions=None, max_len: int = 30,
                   tf_ratio: float = 1.0, *, flags: Optional[Sequence[bool]] = None,
                   hoist: bool = False, aten_lstm: bool = False):
    """Reference: FeaturesCaptioning.decode / forward_sentence
    src/models/features_captioning.py:91-129.

    Returns (sentence [L,B,V] log-probs with row 0 all zero,
             hidden_states [L,1,B,H] with row 0 zero)."""
    B = feats.shape[0]
    H = p[pre + "rnn.weight_hh_l0"].shape[1]
    V = p[pre + "out.weight"].shape[0]
    kw = dict(dtype=feats.dtype)
    h = torch.zeros(B, H, **kw)                                          # :66-75
    c = torch.zeros(B, H, **kw)
    if flags is None:
        flags = teacher_flags(captions, max_len, tf_ratio)
    uv = feats @ p[pre + "attention.U.weight"].t() if hoist else None
    words = torch.full((B,), SOS, dtype=torch.long)                      # :99
    sent, hids = [torch.zeros(B, V, **kw)], [torch.zeros(B, H, **kw)]    # :96-98
    for t in range(1, max_len):                                          # :101
        lp, h, c, _ = decoder_step(p, pre, feats, h, c, words, hoisted_uv=uv, aten_lstm=aten_lstm)
        sent.append(lp)                                                  # :106
        hids.append(h)                                                   # :107
        top1 = lp.detach().argmax(dim=1)                                 # :109
        words = captions[t] if (captions is not None and flags[t - 1]) else top1  # :113-117
    return torch.stack(sent), torch.stack(hids).unsqueeze(1)


def decoder_beam_search(p: Params, pre: str, feats, max_len: int = 30, width: int = 5,
                        alpha: float = 0.0):
    """Reference: FeaturesCaptioning.beam_search_predict
    src/models/features_captioning.py:131-228, restated over [B,width] tensors.

    Per step and per live beam: scores = EOS_mask * log_probs + cum (:166-168),
    divided by ((5+len)^alpha / 6^alpha) for ranking only (:171-180); the
    ``width`` best of the concatenated [B, beams*V] scores are kept (:187-189);
    beam = idx // V, token = idx % V (:192-193); cum carries the UN-normalised
    score (:207).  Ties (only possible between children of a finished beam,
    whose step scores are all zero) are broken towards the lowest flat index;
    the reference's ``argsort`` leaves that order unspecified, so only the
    prefix up to the first EOS of each returned caption is pinned.
    Returns a [B, max_len+2] int64 tensor: SOS then max_len+1 ids of beam 0.
    """
    B = feats.shape[0]
    H = p[pre + "rnn.weight_hh_l0"].shape[1]
    V = p[pre + "out.weight"].shape[0]
    kw = dict(dtype=feats.dtype)
    uv = feats @ p[pre + "attention.U.weight"].t()
    nb = 1
    h = torch.zeros(nb, B, H, **kw)
    c = torch.zeros(nb, B, H, **kw)
    words = torch.full((nb, B), SOS, dtype=torch.long)
    cum = torch.zeros(nb, B, **kw)                                       # log(1) :144-145
    done = torch.zeros(nb, B, dtype=torch.bool)
    length = torch.zeros(nb, B, **kw)     # position of EOS + 1 once finished
    seqs = torch.zeros(nb, B, 0, dtype=torch.long)
    for t in range(max_len + 1):                                         # :149
        sc, rank, hs, cs = [], [], [], []
        for i in range(nb):                                              # :159
            lp, h2, c2, _ = decoder_step(p, pre, feats, h[i], c[i], words[i], hoisted_uv=uv)
            lp = torch.where(done[i].unsqueeze(1), torch.zeros_like(lp), lp) + cum[i].unsqueeze(1)
            clen = torch.where(done[i], length[i], torch.full_like(length[i], t + 1))
            norm = ((5 + clen) ** alpha) / (6 ** alpha)                  # :177
            sc.append(lp)
            rank.append(lp / norm.unsqueeze(1))
            hs.append(h2)
            cs.append(c2)
        sc = torch.cat(sc, dim=1)
        rank = torch.cat(rank, dim=1)                                    # :187
        top = torch.sort(rank, dim=1, descending=True, stable=True)[1][:, :width]  # :189
        bi, wi = top // V, top % V                                       # :192-193
        ar = torch.arange(B)
        hs, cs = torch.stack(hs), torch.stack(cs)
        h = torch.stack([hs[bi[:, k], ar] for k in range(width)])        # :201-206
        c = torch.stack([cs[bi[:, k], ar] for k in range(width)])
        cum = torch.stack([sc[ar, top[:, k]] for k in range(width)])     # :207
        prev_done = torch.stack([done[bi[:, k], ar] for k in range(width)])
        prev_len = torch.stack([length[bi[:, k], ar] for k in range(width)])
        seqs = torch.stack([torch.cat([seqs[bi[:, k], ar], wi[:, k:k + 1]], dim=1)
                            for k in range(width)])                      # :208
        words = wi.t().contiguous()
        newly = (~prev_done) & (words == EOS)
        length = torch.where(prev_done, prev_len, torch.where(newly, torch.full_like(prev_len, t + 1), prev_len))
        done = prev_done | newly
        nb = width
    out = torch.cat([torch.full((B, 1), SOS, dtype=torch.long), seqs[0]], dim=1)  # :227
    return out


# --------------------------------------------------------------------------
# reconstructors
# --------------------------------------------------------------------------
def caption_mask(outputs, captions=None):
    """Reference: build_caption_mask src/models/reconstructor.py:197-206."""
    if captions is None:
        captions = outputs.argmax(dim=2)
    return (captions != PAD) & (captions != EOS)


def global_reconstruct(p: Params, pre: str, hiddens, outputs, captions, *, aten_lstm=False):
    """Reference: GlobalReconstructor.reconstruct
    src/models/reconstructor.py:142-194.  ``hiddens`` [L,1,B,H] ->
    feats_recons [B,L,Fr] (row t=0 zeros)."""
    L, _, B, H = hiddens.shape
    hid = hiddens[:, 0]                                                  # :165-169 (1 layer)
    m = caption_mask(outputs, captions)                                  # :192
    lens = m.sum(dim=0)                                                  # :143
    pooled = (m.unsqueeze(2).to(hid.dtype) * hid).sum(dim=0) / lens.unsqueeze(1)  # :144-148
    Fr = p[pre + "rnn.weight_hh_l0"].shape[1]
    h = torch.zeros(B, Fr, dtype=hid.dtype)
    c = torch.zeros(B, Fr, dtype=hid.dtype)
    rec = [torch.zeros(B, Fr, dtype=hid.dtype)]                          # :175
    for t in range(1, L):                                                # :178
        x = torch.cat([hid[t], pooled], dim=1)                           # :152
        h, c = lstm_cell(x, h, c, p[pre + "rnn.weight_ih_l0"], p[pre + "rnn.weight_hh_l0"],
                         p[pre + "rnn.bias_ih_l0"], p[pre + "rnn.bias_hh_l0"], aten_lstm)
        rec.append(h)                                                    # :183
    return torch.stack(rec).transpose(0, 1)                              # :185


def local_reconstruct(p: Params, pre: str, hiddens, outputs, captions, feat_len: int, *,
                      hoist: bool = False, aten_lstm=False):
    """Reference: LocalReconstructor.reconstruct
    src/models/reconstructor.py:67-97.  -> feats_recons [B,T,Fr]."""
    L, _, B, H = hiddens.shape
    keys = hiddens[:, 0].permute(1, 0, 2)                                # :79-82 -> [B,L,H]
    m = caption_mask(outputs, captions).t()                              # :69, :95
    Fr = p[pre + "rnn.weight_hh_l0"].shape[1]
    h = torch.zeros(B, Fr, dtype=keys.dtype)
    c = torch.zeros(B, Fr, dtype=keys.dtype)
    uk = keys @ p[pre + "attention.U.weight"].t() if hoist else None
    rec = []
    for _ in range(feat_len):                                            # :88
        ctx, _ = soft_attention(p, pre + "attention.", h, keys, mask=m, keys_proj=uk)  # :70
        h, c = lstm_cell(ctx, h, c, p[pre + "rnn.weight_ih_l0"], p[pre + "rnn.weight_hh_l0"],
                         p[pre + "rnn.bias_ih_l0"], p[pre + "rnn.bias_hh_l0"], aten_lstm)  # :73
        rec.append(h)                                                    # :90
    return torch.stack(rec).transpose(0, 1)                              # :91


# --------------------------------------------------------------------------
# losses
# --------------------------------------------------------------------------
def entropy_loss(x, ignore_mask):
    """Reference: EntropyLoss src/losses.py:12-17.  NOTE dim=1 of the
    [L-1,B,V] input is the *batch* axis -- that quirk is part of the value."""
    b = (torch.softmax(x, dim=1) * torch.log_softmax(x, dim=1)).sum(dim=2)
    b = b.masked_fill(ignore_mask, 0.0)
    return -1.0 * b.sum(dim=0).mean()


def global_recon_loss(x, x_rec, keep_mask):
    """Reference: GlobalReconstructionLoss src/losses.py:20-36."""
    xm = x.mean(dim=1)                                                   # :25
    n = keep_mask.sum(dim=0).to(x_rec.dtype).unsqueeze(1)                # :27-28
    k = keep_mask.t().unsqueeze(2).to(x_rec.dtype)                       # :29
    xr = (k * x_rec).sum(dim=1) / n                                      # :34-35
    return torch.nn.functional.mse_loss(xm, xr)                          # :36


def local_recon_loss(x, x_rec):
    """Reference: LocalReconstructionLoss src/losses.py:39-40."""
    return torch.nn.functional.mse_loss(x, x_rec)


def modality_wise_loss(output, captions, audio=None, audio_rec=None, visual=None, visual_rec=None,
                       reg_lambda=0.0, audio_recon_lambda=0.0, visual_recon_lambda=0.0, rec_type="none"):
    """Reference: ModalityWiseReconstructionLoss src/losses.py:86-126.
    Returns (loss, ce, entropy, audio_rec_loss, visual_rec_loss)."""
    def rec(feats, feats_rec):
        if feats_rec is None or rec_type not in ("global", "local"):    # :100-101
            return torch.zeros((), dtype=output.dtype)
        if rec_type == "global":
            return global_recon_loss(feats, feats_rec, captions != PAD)  # :104
        return local_recon_loss(feats, feats_rec)                        # :106
    V = output.shape[2]
    ce = torch.nn.functional.nll_loss(output[1:].reshape(-1, V), captions[1:].reshape(-1),
                                      ignore_index=PAD)                  # :112
    ent = entropy_loss(output[1:], captions[1:] == PAD)                  # :115
    a, v = rec(audio, audio_rec), rec(visual, visual_rec)                # :118-119
    loss = ce + reg_lambda * ent + audio_recon_lambda * a + visual_recon_lambda * v  # :122-124
    return loss, ce, ent, a, v


# --------------------------------------------------------------------------
# model wrappers
# --------------------------------------------------------------------------
def _recon(p, pre, rec_type, hiddens, outputs, captions, T, **kw):
    if rec_type == "global":
        kw.pop("hoist", None)
        return global_reconstruct(p, pre, hiddens, outputs, captions, **kw)
    if rec_type == "local":
        return local_reconstruct(p, pre, hiddens, outputs, captions, T, **kw)
    return None


def av_forward(p: Params, audio, visual, captions, tf_ratio=1.0, rec_type="none", *,
               flags=None, hoist=False, aten_lstm=False):
    """Reference: AVCaptioning.forward src/models/captioning.py:108-128
    (early fusion: cat([audio, visual], -1), audio first)."""
    feats = torch.cat([audio, visual], dim=-1)                           # :109
    out, hid = decoder_decode(p, "decoder.", feats, captions, captions.shape[0], tf_ratio,
                              flags=flags, hoist=hoist, aten_lstm=aten_lstm)
    rec = _recon(p, "reconstructor.", rec_type, hid, out, captions, feats.shape[1],
                 hoist=hoist, aten_lstm=aten_lstm)
    if rec is None:
        return out, None, None
    Fa = audio.shape[2]
    return out, rec[:, :, :Fa], rec[:, :, Fa:]                           # :125-126


def av_dual_forward(p: Params, audio, visual, captions, tf_ratio=1.0, rec_type="none", *,
                    flags_v=None, flags_a=None, hoist=False, aten_lstm=False):
    """Reference: AVCaptioningDual.forward src/models/captioning.py:223-258
    (late fusion: the two log-prob tensors are summed, :260-264).  The visual
    decoder runs first (and draws its RNG numbers first)."""
    L = captions.shape[0]
    vo, vh = decoder_decode(p, "v_decoder.", visual, captions, L, tf_ratio, flags=flags_v,
                            hoist=hoist, aten_lstm=aten_lstm)
    ao, ah = decoder_decode(p, "a_decoder.", audio, captions, L, tf_ratio, flags=flags_a,
                            hoist=hoist, aten_lstm=aten_lstm)
    out = ao + vo
    ar = _recon(p, "a_reconstructor.", rec_type, ah, ao, captions, audio.shape[1],
                hoist=hoist, aten_lstm=aten_lstm)
    vr = _recon(p, "v_reconstructor.", rec_type, vh, vo, captions, visual.shape[1],
                hoist=hoist, aten_lstm=aten_lstm)
    return out, ar, vr


def av_greedy_ids(p: Params, audio, visual, max_len=30, *, hoist=True):
    """Reference: AVCaptioning.predict(mode="direct")
    src/models/captioning.py:131-141 -> ids [B,max_len] (column 0 is the
    argmax of an all-zero row, i.e. 0)."""
    feats = torch.cat([audio, visual], dim=-1)
    out, _ = decoder_decode(p, "decoder.", feats, None, max_len, hoist=hoist)
    return out.argmax(2).transpose(1, 0)


def av_dual_greedy_ids(p: Params, audio, visual, max_len=30, *, hoist=True):
    """Reference: AVCaptioningDual.predict(mode="direct")
    src/models/captioning.py:279-287."""
    vo, _ = decoder_decode(p, "v_decoder.", visual, None, max_len, hoist=hoist)
    ao, _ = decoder_decode(p, "a_decoder.", audio, None, max_len, hoist=hoist)
    return (ao + vo).argmax(2).transpose(1, 0)


def decode_indexes(itos, ids) -> str:
    """Reference: Vocabulary.decode_indexes src/get_loader.py:79-89."""
    words = []
    for i in ids:
        i = int(i)
        if i == EOS:
            break
        words.append(itos[i])
    return " ".join(words)


# --------------------------------------------------------------------------
# parameter construction (reference default init)
# --------------------------------------------------------------------------
def init_decoder_params(pre: str, F: int, V: int, H=512, E=300, A=256, gen=None, dtype=torch.float32) -> Params:
    """Same shapes/names as FeaturesCaptioning.state_dict()
    (features_captioning.py:36-56).  Embedding N(0,1); Linear/LSTM
    U(+-1/sqrt(fan)); attention.b = 1 (temporal_attention.py:16)."""
    def u(*shape, fan):
        k = 1.0 / math.sqrt(fan)
        return ((torch.rand(*shape, generator=gen, dtype=torch.float64) * 2 - 1) * k).to(dtype)
    p = {
        pre + "embedding.weight": torch.randn(V, E, generator=gen, dtype=torch.float64).to(dtype),
        pre + "attention.b": torch.ones(A, dtype=dtype),
        pre + "attention.W.weight": u(A, H, fan=H),
        pre + "attention.U.weight": u(A, F, fan=F),
        pre + "attention.w.weight": u(1, A, fan=A),
        pre + "rnn.weight_ih_l0": u(4 * H, E + F, fan=H),
        pre + "rnn.weight_hh_l0": u(4 * H, H, fan=H),
        pre + "rnn.bias_ih_l0": u(4 * H, fan=H),
        pre + "rnn.bias_hh_l0": u(4 * H, fan=H),
        pre + "out.weight": u(V, H, fan=H),
        pre + "out.bias": u(V, fan=H),
    }
    return p


def init_recon_params(pre: str, kind: str, dec_H: int, Fr: int, A=256, gen=None, dtype=torch.float32) -> Params:
    """Same shapes/names as Global/LocalReconstructor.state_dict()
    (reconstructor.py:34-46, :123-129)."""
    def u(*shape, fan):
        k = 1.0 / math.sqrt(fan)
        return ((torch.rand(*shape, generator=gen, dtype=torch.float64) * 2 - 1) * k).to(dtype)
    In = 2 * dec_H if kind == "global" else dec_H
    p = {
        pre + "rnn.weight_ih_l0": u(4 * Fr, In, fan=Fr),
        pre + "rnn.weight_hh_l0": u(4 * Fr, Fr, fan=Fr),
        pre + "rnn.bias_ih_l0": u(4 * Fr, fan=Fr),
        pre + "rnn.bias_hh_l0": u(4 * Fr, fan=Fr),
    }
    if kind == "local":
        p.update({
            pre + "attention.b": torch.ones(A, dtype=dtype),
            pre + "attention.W.weight": u(A, Fr, fan=Fr),
            pre + "attention.U.weight": u(A, dec_H, fan=dec_H),
            pre + "attention.w.weight": u(1, A, fan=A),
        })
    return p


# --------------------------------------------------------------------------
# synthetic inputs
# --------------------------------------------------------------------------
def synth_batch(B: int, T: int, L: int, V: int, Fa=128, Fv=2048, seed=1, min_frames=4, min_cap=8):
    """MSVD / MSR-VTT-shaped synthetic batch in the loader's layout
    (CustomCollateAV src/get_loader.py:403-413): audio [B,T,Fa] f32 with
    integer values 0..255, visual [B,T,Fv] f32 = relu(randn) (max ~5; scaled so
    the max is ~48 like Inception pool features), zero-padded trailing frames,
    captions [L,B] i64 with SOS first, EOS at len-1 and PAD after."""
    g = torch.Generator().manual_seed(seed)
    audio = torch.randint(0, 256, (B, T, Fa), generator=g).float()
    visual = torch.relu(torch.randn(B, T, Fv, generator=g)) * 10.0
    nfr = torch.randint(min(min_frames, T), T + 1, (B,), generator=g)
    fr_mask = (torch.arange(T).unsqueeze(0) < nfr.unsqueeze(1)).unsqueeze(2)
    audio, visual = audio * fr_mask, visual * fr_mask
    lens = torch.randint(min(min_cap, L), L + 1, (B,), generator=g)
    lens[0] = L  # at least one full-length caption, as pad_sequence guarantees
    caps = torch.randint(4, V, (L, B), generator=g)
    pos = torch.arange(L).unsqueeze(1)
    caps = torch.where(pos == 0, torch.full_like(caps, SOS), caps)
    caps = torch.where(pos == (lens - 1).unsqueeze(0), torch.full_like(caps, EOS), caps)
    caps = torch.where(pos >= lens.unsqueeze(0), torch.full_like(caps, PAD), caps)
    return audio, visual, caps
