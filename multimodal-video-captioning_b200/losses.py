"""Drop-in for the hot-path half of the reference's `losses` module (src/losses.py:12-137):
same function names and signatures, evaluated by the fused CUDA loss kernels of libmvc_b200
(forward value and gradient).  NLPScore (losses.py:140-160: BLEU / METEOR / ROUGE-L / CIDEr over pycocoevalcap) is
provided too, so that `from losses import ..., NLPScore` (train.py:12) resolves: BLEU, ROUGE-L and CIDEr come from
salstm/nlp_score.py (pinned against the reference's scorers); METEOR, a Java program, is delegated to the reference's
scorer found further down sys.path when it and `java` exist."""
import importlib.util
import os
import sys
from functools import partial

import torch

from salstm import functional as Fn


def ModalityWiseReconstructionLoss(output, captions, audio_features=None, audio_features_recons=None,
                                   visual_features=None, visual_features_recons=None, reg_lambda=0,
                                   audio_recon_lambda=0, visual_recon_lambda=0, rec_type="none"):
    """losses.py:86-126 -> (loss, ce, entropy, audio_rec_loss, visual_rec_loss)."""
    return Fn.ModalityLossFn.apply(output, captions, audio_features, audio_features_recons, visual_features,
                                   visual_features_recons, float(reg_lambda), float(audio_recon_lambda),
                                   float(visual_recon_lambda), rec_type)


def ModalityWiseReconstructionLossBuilder(reg_lambda, audio_recon_lambda, visual_recon_lambda, rec_type="none"):
    """losses.py:129-137."""
    assert rec_type in ["none", "global", "local"], "Wrong mode specified, must be one of ['none', 'global', 'local']"
    return partial(ModalityWiseReconstructionLoss, reg_lambda=reg_lambda, audio_recon_lambda=audio_recon_lambda,
                   visual_recon_lambda=visual_recon_lambda, rec_type=rec_type)


def TotalReconstructionLoss(output, captions, features=None, features_recons=None, reg_lambda=0, recon_lambda=0,
                            reconstruction_type="global"):
    """losses.py:43-69 (single-stream twin) -> (loss, ce, entropy, rec_loss)."""
    loss, ce, ent, _, rec = Fn.ModalityLossFn.apply(output, captions, None, None, features, features_recons,
                                                    float(reg_lambda), 0.0, float(recon_lambda), reconstruction_type)
    if features_recons is None or reconstruction_type not in ("global", "local"):
        loss, rec = loss.reshape(1), rec.reshape(1)         # the reference's torch.zeros(1) term broadcasts (:58-66)
    return loss, ce, ent, rec


def ReconstructionLossBuilder(reg_lambda, recon_lambda, reconstruction_type):
    """losses.py:72-83."""
    assert reconstruction_type in ["none", "global", "local"], \
        "Wrong mode specified, must be one of ['none', 'global', 'local']"
    return partial(TotalReconstructionLoss, reg_lambda=reg_lambda, recon_lambda=recon_lambda,
                   reconstruction_type=reconstruction_type)


def EntropyLoss(x, ignore_mask):
    """losses.py:12-17 on an [S,B,V] log-prob tensor (softmax over the batch axis, quirk kept)."""
    S, B, V = x.shape
    pad = torch.zeros(1, B, V, device=x.device, dtype=x.dtype)
    caps = torch.cat([torch.ones(1, B, device=x.device, dtype=torch.int64), (~ignore_mask).to(torch.int64)], 0)
    _, _, ent, _, _ = Fn.ModalityLossFn.apply(torch.cat([pad, x.detach()], 0), caps, None, None, None, None, 0.0, 0.0,
                                              0.0, "none")
    return ent


_REF_LOSSES = None


def _reference_losses(required: bool = False):
    """The reference's own losses.py: the next `losses.py` on sys.path after this one (the launcher and
    INTEGRATION.md put the reference's src/ and repository root behind this package); None when absent."""
    global _REF_LOSSES
    if _REF_LOSSES is None:
        here = os.path.dirname(os.path.abspath(__file__))
        for d in sys.path:
            cand = os.path.join(d or os.getcwd(), "losses.py")
            if os.path.isfile(cand) and os.path.abspath(os.path.dirname(cand)) != here:
                root = os.path.dirname(os.path.dirname(os.path.abspath(cand)))       # <reference>/ holds pycocoevalcap/
                if os.path.isdir(os.path.join(root, "pycocoevalcap")) and root not in sys.path:
                    sys.path.append(root)
                try:
                    spec = importlib.util.spec_from_file_location("_mvc_reference_losses", cand)
                    mod = importlib.util.module_from_spec(spec)
                    spec.loader.exec_module(mod)
                    _REF_LOSSES = mod
                except Exception:
                    _REF_LOSSES = None
                break
        if _REF_LOSSES is None and required:
            raise ImportError("the reference's src/losses.py is not on sys.path behind multimodal-video-captioning_b200/")
    return _REF_LOSSES


def NLPScore(ref, hypo):
    """losses.py:140-160 -> {"Bleu_1".."Bleu_4", "METEOR", "ROUGE_L", "CIDEr"} for dicts {id: [sentence, ...]}."""
    from salstm.nlp_score import nlp_score
    return nlp_score(ref, hypo, _reference_losses())
