"""Drop-in for the hot-path half of the reference's `losses` module (src/losses.py:12-137):
same function names and signatures, evaluated by the fused CUDA loss kernels of libmvc_b200
(forward value and gradient).  NLPScore (string metrics, losses.py:140-160) is out of scope and
stays the reference's."""
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
