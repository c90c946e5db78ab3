"""Host-side mirror of the reference's nn.Module surface for the hot path.

Same class names, constructor kwargs, parameter names (``state_dict`` keys) and
method signatures as ``src/models/{temporal_attention,features_captioning,
reconstructor,captioning}.py`` of the reference, so ``src/train.py`` and
``notebooks/predict_captions.ipynb`` run unchanged with this package earlier on
``sys.path``.  The modules hold parameters and shapes only; all arithmetic is in
libmvc_b200 (CUDA, sm_100a).  There is no CPU path: compute on a non-CUDA tensor
raises RuntimeError; GRU / bidirectional / multi-layer configurations (never
enabled by any reference config) raise NotImplementedError at construction.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch
import torch.nn as nn

from . import cabi
from . import functional as Fn

_DEFAULT_PRECISION = os.environ.get("MVC_B200_PRECISION", "fp32")


def _unsupported(rnn_type, num_layers, bidirectional):
    if rnn_type != "LSTM" or num_layers != 1 or bidirectional:
        raise NotImplementedError(
            "mvc_b200 implements the configuration every reference experiment uses (rnn_type='LSTM', "
            f"rnn_num_layers=1, rnn_bidirectional=False); got {rnn_type!r}, {num_layers}, {bidirectional}. "
            "There is no CPU/ATen fallback for other settings.")


class TemporalAttention(nn.Module):
    """Additive attention parameters (temporal_attention.py:7-17) + a standalone forward."""

    def __init__(self, hidden_size, feature_size, bottleneck_size):
        super().__init__()
        self.hidden_size, self.feature_size, self.bottleneck_size = hidden_size, feature_size, bottleneck_size
        self.W = nn.Linear(hidden_size, bottleneck_size, bias=False)
        self.U = nn.Linear(feature_size, bottleneck_size, bias=False)
        self.b = nn.Parameter(torch.ones(bottleneck_size), requires_grad=True)
        self.w = nn.Linear(bottleneck_size, 1, bias=False)

    def forward(self, hidden, feats, masks=None):
        """(attn_feats [B,F], weights [B,T,1]); temporal_attention.py:19-33.  Differentiable w.r.t. hidden,
        feats and the four parameters (Fn.SoftAttentionFn); inside the decoder / reconstructor time loops the
        attention runs fused in the step kernels instead."""
        ctx, alpha = Fn.SoftAttentionFn.apply(hidden, feats, masks, self.W.weight, self.U.weight, self.b,
                                              self.w.weight)
        return ctx, alpha.unsqueeze(2)


class FeaturesCaptioning(nn.Module):
    """SA-LSTM caption decoder; features_captioning.py:9-228."""
    # class-level default: modules unpickled from a checkpoint the REFERENCE saved (torch.save(model),
    # train.py:162-173) carry no `precision` in their __dict__ and run the exact fp32 path
    precision = _DEFAULT_PRECISION

    def __init__(self, in_feature_size, output_size, rnn_type="LSTM", rnn_num_layers=1, rnn_bidirectional=False,
                 rnn_hidden_size=128, rnn_dropout=0.5, embedding_size=128, attn_size=128, device="cpu",
                 precision=None, **args):
        super().__init__()
        _unsupported(rnn_type, rnn_num_layers, rnn_bidirectional)
        self.rnn_type, self.num_layers, self.num_directions = rnn_type, 1, 1
        self.feature_size, self.embedding_size = in_feature_size, embedding_size
        self.hidden_size, self.attn_size, self.output_size = rnn_hidden_size, attn_size, output_size
        self.rnn_dropout_p = 0          # dropout is a no-op for one layer (features_captioning.py:33)
        self.device = device
        self.precision = precision or _DEFAULT_PRECISION
        self.embedding = nn.Embedding(output_size, embedding_size)
        self.attention = TemporalAttention(rnn_hidden_size, in_feature_size, attn_size)
        # nn.LSTM is used as the parameter container only: same names / shapes / default init as the reference
        self.rnn = nn.LSTM(input_size=embedding_size + in_feature_size, hidden_size=rnn_hidden_size, num_layers=1)
        self.out = nn.Linear(rnn_hidden_size, output_size)

    # ---- helpers
    def _params(self):
        sd = dict(self.named_parameters())
        return [sd[k] for k in Fn.DEC_PARAM_ORDER]

    def _dims(self, B, T, L):
        return (B, T, self.feature_size, self.hidden_size, self.embedding_size, self.attn_size, self.output_size, L,
                cabi.precision_id(self.precision))

    @staticmethod
    def _split(features):
        """features is either one [B,T,F] tensor or an (audio, visual) pair fused in-kernel."""
        if isinstance(features, (tuple, list)):
            return features[0], features[1]
        return None, features

    def _init_hidden(self, batch_size):
        dev = self.out.weight.device
        h, c = (torch.zeros(1, batch_size, self.hidden_size, device=dev) for _ in range(2))
        h._mvc_zero_state = True            # forward_sentence: this state takes the fused time loop
        return h, c

    # ---- reference API
    def decode(self, features, captions=None, max_caption_len=30, teacher_forcing_ratio=1):
        """-> (outputs [L,B,V] log-probs, decoder_hiddens [L,1,B,H]); features_captioning.py:121-129."""
        a, v = self._split(features)
        ref = v if v is not None else a
        B, T = ref.shape[0], ref.shape[1]
        flags = Fn.teacher_flags(captions, max_caption_len, teacher_forcing_ratio)
        return Fn.DecoderFn.apply(self._dims(B, T, max_caption_len), flags, a, v, captions, *self._params())

    def forward_sentence(self, features, captions, hidden, max_caption_len=30, teacher_forcing_ratio=1):
        """features_captioning.py:91-119.  The zero state of _init_hidden (the only state the reference ever
        passes, :124-125) takes the fused time-loop kernels; any other initial state is stepped word by word
        through forward_word (same arithmetic and RNG draws, differentiable, one launch chain per word)."""
        if hidden is None or getattr(hidden[0], "_mvc_zero_state", False):
            return self.decode(features, captions, max_caption_len, teacher_forcing_ratio)
        feats = features if not isinstance(features, (tuple, list)) else torch.cat(list(features), -1)
        B = feats.shape[0]
        dev = feats.device
        zeros_o = torch.zeros(B, self.output_size, device=dev)
        zeros_h = torch.zeros(1, B, self.hidden_size, device=dev)
        words = torch.full((1, B), 1, dtype=torch.int64, device=dev)                 # <SOS>, :101
        sentence, hiddens = [zeros_o], [zeros_h]
        for t in range(1, max_caption_len):
            logp, hidden, _ = self.forward_word(feats, hidden, words)
            sentence.append(logp)
            hiddens.append(hidden[0])
            teacher = captions is not None and bool(torch.rand(1) < teacher_forcing_ratio)   # :113-116
            words = (captions[t] if teacher else logp.detach().argmax(1)).reshape(1, B)
        return torch.stack(sentence), torch.stack(hiddens)

    forward = decode

    def forward_word(self, features, hidden, previous_words):
        """One decoder step from an arbitrary state: (log_probs [B,V], (h,c), attn_weights [B,T,1]);
        features_captioning.py:77-89.  Built from the fp32 block kernels as two autograd nodes
        (Fn.SoftAttentionFn, Fn.WordStepFn), so stepping the decoder word by word is differentiable
        like the reference's."""
        h0, c0 = hidden
        feats = features if not isinstance(features, (tuple, list)) else torch.cat(list(features), -1)
        ctx, alpha = self.attention(h0[-1], feats)
        logp, h1, c1 = Fn.WordStepFn.apply(previous_words, ctx, h0[-1], c0[-1], self.embedding.weight,
                                           self.rnn.weight_ih_l0, self.rnn.weight_hh_l0, self.rnn.bias_ih_l0,
                                           self.rnn.bias_hh_l0, self.out.weight, self.out.bias)
        return logp, (h1.unsqueeze(0), c1.unsqueeze(0)), alpha

    @torch.no_grad()
    def greedy_ids(self, features, max_caption_len=30):
        """ids [B, L] of the free-running decode (== decode(features).argmax(2).T) without
        materialising [L,B,V]."""
        a, v = self._split(features)
        ref = v if v is not None else a
        return Fn.decoder_greedy(self._dims(ref.shape[0], ref.shape[1], max_caption_len), a, v, self._params())

    @torch.no_grad()
    def beam_search_predict(self, features, vocab, max_caption_len=30, beam_alpha=0, beam_width=5):
        """-> list[B] of [SOS] + (max_caption_len+1) ids; features_captioning.py:131-228.
        `vocab` is accepted for signature parity; SOS/EOS ids are the loader's fixed 1/2."""
        a, v = self._split(features)
        ref = v if v is not None else a
        if vocab is not None and hasattr(vocab, "stoi"):
            assert vocab.stoi.get("<SOS>", 1) == 1 and vocab.stoi.get("<EOS>", 2) == 2
        ids = Fn.decoder_beam(self._dims(ref.shape[0], ref.shape[1], max_caption_len), a, v, self._params(),
                              beam_width, float(beam_alpha))
        return ids.cpu().tolist()


class _ReconBase(nn.Module):
    precision = _DEFAULT_PRECISION      # see FeaturesCaptioning.precision

    def _dims(self, B, L, T):
        return (B, L, self.decoder_size, self.hidden_size, getattr(self, "attn_size", 0) or 0, T,
                cabi.precision_id(self.precision))

    def _params(self):
        sd = dict(self.named_parameters())
        return [sd[k] for k in Fn.REC_PARAM_ORDER if k in sd]


class GlobalReconstructor(_ReconBase):
    """reconstructor.py:100-194."""

    def __init__(self, decoder_size, hidden_size, rnn_type="LSTM", rnn_num_layers=1, rnn_bidirectional=False,
                 rnn_dropout=0.5, device="cpu", precision=None, **args):
        super().__init__()
        _unsupported(rnn_type, rnn_num_layers, rnn_bidirectional)
        self._type = "global"
        self.rnn_type, self.num_layers, self.num_directions = rnn_type, 1, 1
        self.decoder_size, self.hidden_size, self.device = decoder_size, hidden_size, device
        self.rnn_dropout_p = 0
        self.precision = precision or _DEFAULT_PRECISION
        self.rnn = nn.LSTM(input_size=decoder_size * 2, hidden_size=hidden_size, num_layers=1)

    def reconstruct(self, decoder_hiddens, outputs, captions, target_feature_length=None):
        """decoder_hiddens [L,1,B,H] -> feats_recons [B,L,Fr]; reconstructor.py:187-194."""
        L, _, B, _ = decoder_hiddens.shape
        mask = Fn.caption_mask(outputs, captions)
        return Fn.GlobalReconFn.apply(self._dims(B, L, 0), decoder_hiddens, mask, *self._params())


class LocalReconstructor(_ReconBase):
    """reconstructor.py:9-97."""

    def __init__(self, decoder_size, hidden_size, rnn_type="LSTM", rnn_num_layers=1, rnn_bidirectional=False,
                 rnn_dropout=0.5, attn_size=128, device="cpu", precision=None, **args):
        super().__init__()
        _unsupported(rnn_type, rnn_num_layers, rnn_bidirectional)
        self._type = "local"
        self.rnn_type, self.num_layers, self.num_directions = rnn_type, 1, 1
        self.decoder_size, self.hidden_size, self.attn_size, self.device = decoder_size, hidden_size, attn_size, device
        self.rnn_dropout_p = 0
        self.precision = precision or _DEFAULT_PRECISION
        self.rnn = nn.LSTM(input_size=decoder_size, hidden_size=hidden_size, num_layers=1)
        self.attention = TemporalAttention(hidden_size=hidden_size, feature_size=decoder_size, bottleneck_size=attn_size)

    def reconstruct(self, decoder_hiddens, outputs, captions, target_feature_length):
        """decoder_hiddens [L,1,B,H] -> feats_recons [B,T,Fr]; reconstructor.py:94-97."""
        L, _, B, _ = decoder_hiddens.shape
        mask = Fn.caption_mask(outputs, captions)
        return Fn.LocalReconFn.apply(self._dims(B, L, int(target_feature_length)), decoder_hiddens, mask, *self._params())


def build_caption_mask(outputs, captions=None):
    """bool [L,B]; reconstructor.py:197-206."""
    return Fn.caption_mask(outputs, captions).bool()


def decode_batch(vocab, ids):
    """[vocab.decode_indexes(row[1:]) for row in ids] (get_loader.py:79-89: words up to the first <EOS>, joined by
    spaces) for a whole [B, L] id matrix at once: one table lookup over the matrix and one first-EOS search instead of a
    Python loop with a dict lookup and an isinstance check per token.  Vocabularies without a dense
    `itos` table fall back to the vocabulary's own method."""
    import numpy as np
    arr = np.asarray(ids, dtype=np.int64)
    itos = getattr(vocab, "itos", None)
    if arr.ndim != 2 or not isinstance(itos, dict) or len(itos) == 0 or arr.size == 0 or \
            int(arr[:, 1:].max(initial=0)) >= len(itos) or int(arr[:, 1:].min(initial=0)) < 0:
        return [vocab.decode_indexes(o[1:]) for o in (ids.tolist() if hasattr(ids, "tolist") else ids)]
    table = getattr(vocab, "_mvc_itos_table", None)
    if table is None or len(table) != len(itos):
        try:
            table = np.array([itos[i] for i in range(len(itos))], dtype=object)
        except KeyError:
            return [vocab.decode_indexes(o[1:]) for o in arr.tolist()]
        try:
            vocab._mvc_itos_table = table
        except Exception:
            pass
    body = arr[:, 1:]
    is_eos = body == 2
    first = np.where(is_eos.any(1), is_eos.argmax(1), body.shape[1])
    words = table[body]
    return [" ".join(words[b, :first[b]]) for b in range(body.shape[0])]


# --------------------------------------------------------------------------- wrappers (captioning.py)
DECODER_CONFIG = {"rnn_type": "LSTM", "rnn_num_layers": 1, "rnn_bidirectional": False, "rnn_hidden_size": 512,
                  "rnn_dropout": 0.0, "in_feature_size": 2048 + 128, "embedding_size": 300, "attn_size": 256,
                  "output_size": 1024}
RECONSTRUCTOR_CONFIG = {"type": "global", "rnn_type": "LSTM", "rnn_num_layers": 1, "rnn_bidirectional": False,
                        "hidden_size": 2048 + 128, "rnn_dropout": 0.5, "decoder_size": 512, "attn_size": 256}
VISUAL_DECODER_CONFIG = dict(DECODER_CONFIG, in_feature_size=2048)
AUDIO_DECODER_CONFIG = dict(DECODER_CONFIG, in_feature_size=128, output_size=512)


def _make_recon(kind, rec_config, device, precision):
    if kind == "global":
        return GlobalReconstructor(**rec_config, device=device, precision=precision).to(device)
    if kind == "local":
        return LocalReconstructor(**rec_config, device=device, precision=precision).to(device)
    return None


class AVCaptioning(nn.Module):
    """Early-fusion model; captioning.py:58-144."""

    def __init__(self, vocab, teacher_forcing_ratio=0.0, reconstructor_type="none", device="cpu",
                 normalize_inputs=False, precision=None):
        super().__init__()
        self.vocab, self.vocab_size = vocab, len(vocab)
        self.teacher_forcing_ratio, self.normalize_inputs = teacher_forcing_ratio, normalize_inputs
        config = dict(DECODER_CONFIG, output_size=self.vocab_size)
        rec_config = dict(RECONSTRUCTOR_CONFIG, decoder_size=config["rnn_hidden_size"],
                          hidden_size=config["in_feature_size"], type=reconstructor_type)
        self.decoder = FeaturesCaptioning(**config, device=device, precision=precision).to(device)
        self.reconstructor = _make_recon(reconstructor_type, rec_config, device, precision)
        self.reconstructor_type = reconstructor_type

    def set_precision(self, precision):
        for m in self.modules():
            if hasattr(m, "precision"):
                m.precision = precision
        return self

    def forward(self, audio_features, visual_features, captions, teacher_forcing_ratio=None):
        """-> (outputs [L,B,V], audio_recons, visual_recons); captioning.py:108-128.  The early-fusion
        cat([audio, visual], -1) (:109) happens inside the decoder's first kernel."""
        tf = teacher_forcing_ratio if teacher_forcing_ratio is not None else self.teacher_forcing_ratio
        outputs, rnn_hiddens = self.decoder.decode((audio_features, visual_features), captions,
                                                   max_caption_len=captions.shape[0], teacher_forcing_ratio=tf)
        if self.reconstructor is None:
            return outputs, None, None
        rec = self.reconstructor.reconstruct(rnn_hiddens, outputs, captions, visual_features.shape[1])
        a_rec, v_rec = Fn.SplitFeaturesFn.apply(rec, audio_features.shape[2])     # :125-126
        return outputs, a_rec, v_rec

    @torch.no_grad()
    def predict_ids(self, audio_features, visual_features, max_caption_len=30, mode="direct", beam_alpha=0,
                    beam_width=5):
        feats = (audio_features, visual_features)
        if mode == "beam":
            return self.decoder.beam_search_predict(feats, self.vocab, max_caption_len, beam_alpha, beam_width)
        if mode == "direct":
            return self.decoder.greedy_ids(feats, max_caption_len).cpu().tolist()
        raise ValueError(f"unknown mode {mode!r}")

    def predict(self, audio_features, visual_features, max_caption_len=30, mode="direct", beam_alpha=0, beam_width=5):
        """-> list[str]; captioning.py:131-144."""
        ids = self.predict_ids(audio_features, visual_features, max_caption_len, mode, beam_alpha, beam_width)
        return decode_batch(self.vocab, ids)


class AVCaptioningDual(nn.Module):
    """Late-fusion model (two decoders, log-probs summed); captioning.py:147-291."""

    def __init__(self, vocab, teacher_forcing_ratio=0.0, reconstructor_type="none", device="cpu",
                 normalize_inputs=False, precision=None):
        super().__init__()
        self.vocab, self.vocab_size = vocab, len(vocab)
        self.teacher_forcing_ratio, self.normalize_inputs = teacher_forcing_ratio, normalize_inputs
        v_config = dict(VISUAL_DECODER_CONFIG, output_size=self.vocab_size)
        a_config = dict(AUDIO_DECODER_CONFIG, output_size=self.vocab_size)
        v_rec = dict(RECONSTRUCTOR_CONFIG, decoder_size=v_config["rnn_hidden_size"],
                     hidden_size=v_config["in_feature_size"], type=reconstructor_type)
        a_rec = dict(RECONSTRUCTOR_CONFIG, decoder_size=a_config["rnn_hidden_size"],
                     hidden_size=a_config["in_feature_size"], type=reconstructor_type)
        self.v_decoder = FeaturesCaptioning(**v_config, device=device, precision=precision).to(device)
        self.a_decoder = FeaturesCaptioning(**a_config, device=device, precision=precision).to(device)
        # kept for state_dict / optimizer parity: the reference builds it and never uses it (captioning.py:185)
        self.output_fc = nn.Linear(a_config["output_size"] + v_config["output_size"], self.vocab_size)
        self.v_reconstructor = _make_recon(reconstructor_type, v_rec, device, precision)
        self.a_reconstructor = _make_recon(reconstructor_type, a_rec, device, precision)
        self.reconstructor_type = reconstructor_type

    set_precision = AVCaptioning.set_precision

    def _feature_fusion(self, a_outputs, v_outputs):
        return a_outputs + v_outputs                                           # captioning.py:260-264

    def forward(self, audio_features, visual_features, captions, teacher_forcing_ratio=None):
        tf = teacher_forcing_ratio if teacher_forcing_ratio is not None else self.teacher_forcing_ratio
        L = captions.shape[0]
        v_out, v_hid = self.v_decoder.decode(visual_features, captions, max_caption_len=L, teacher_forcing_ratio=tf)
        a_out, a_hid = self.a_decoder.decode(audio_features, captions, max_caption_len=L, teacher_forcing_ratio=tf)
        outputs = self._feature_fusion(a_out, v_out)
        a_rec = None if self.a_reconstructor is None else \
            self.a_reconstructor.reconstruct(a_hid, a_out, captions, audio_features.shape[1])
        v_rec = None if self.v_reconstructor is None else \
            self.v_reconstructor.reconstruct(v_hid, v_out, captions, visual_features.shape[1])
        return outputs, a_rec, v_rec

    @torch.no_grad()
    def predict_ids(self, audio_features, visual_features, max_caption_len=30, mode="direct", beam_alpha=0,
                    beam_width=5):
        if mode != "direct":
            # the reference's Dual beam mode raises UnboundLocalError (captioning.py:269-277, "FIXME: not implemented")
            raise NotImplementedError("AVCaptioningDual.predict(mode='beam') is not implemented in the reference either")
        lib = cabi.lib()
        v_out, _ = self.v_decoder.decode(visual_features, None, max_caption_len=max_caption_len)
        a_out, _ = self.a_decoder.decode(audio_features, None, max_caption_len=max_caption_len)
        L, B, V = v_out.shape
        ids = torch.empty(L, B, device=v_out.device, dtype=torch.int64)
        cabi.check(lib.mvc_argmax_rows(cabi.ptr(a_out), cabi.ptr(v_out), L * B, V, cabi.ptr(ids), cabi.stream_ptr()))
        return ids.t().cpu().tolist()

    def predict(self, audio_features, visual_features, max_caption_len=30, mode="direct", beam_alpha=0, beam_width=5):
        ids = self.predict_ids(audio_features, visual_features, max_caption_len, mode, beam_alpha, beam_width)
        return decode_batch(self.vocab, ids)


# Pickle identity.  torch.save(model) (train.py:162-173) records each class as module path + name; the reference's
# classes live in models.captioning / models.features_captioning / models.reconstructor / models.temporal_attention.
# Reporting the same paths makes (i) whole-module pickles written by the reference load into these classes and
# (ii) pickles written from here resolve through `models.*` wherever they are loaded.
for _cls, _mod in ((TemporalAttention, "models.temporal_attention"), (FeaturesCaptioning, "models.features_captioning"),
                   (GlobalReconstructor, "models.reconstructor"), (LocalReconstructor, "models.reconstructor"),
                   (AVCaptioning, "models.captioning"), (AVCaptioningDual, "models.captioning")):
    _cls.__module__ = _mod
del _cls, _mod
