"""Round-2 fixtures from the UNMODIFIED reference (build container only; needs /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python tools/make_golden_r2.py

  tests/golden/ref_decoder_tiny.pt   torch.save(<reference FeaturesCaptioning>) -- a whole-module pickle as
                                     train.py:162-173 writes them (class path models.features_captioning.*),
                                     small enough to commit; plus a GlobalReconstructor / LocalReconstructor pickle
  tests/golden/ref_pickle_expect.npz inputs + the reference's outputs for the unpickled modules (teacher-forced
                                     log-probs, greedy ids, beam ids, reconstructions)
  tests/golden/word_step_small.npz   forward_word / forward_sentence from a NON-zero hidden state, with gradients
                                     of every parameter and of the inputs; stand-alone TemporalAttention.forward
                                     gradients (masked)
  tests/golden/total_loss_small.npz  TotalReconstructionLoss (the single-stream twin, losses.py:43-69): values and
                                     gradients for none / global / local
"""
import os
import sys

os.environ.setdefault("PYTHONDONTWRITEBYTECODE", "1")
sys.dont_write_bytecode = True
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path[:0] = ["/root/reference/src", "/root/reference"]

import numpy as np
import torch

from oracle import salstm_oracle as O

import losses as ref_losses  # noqa: E402  (reference)
from models import FeaturesCaptioning  # noqa: E402
from models.reconstructor import GlobalReconstructor, LocalReconstructor  # noqa: E402
from models.temporal_attention import TemporalAttention  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def npy(d):
    return {k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in d.items()}


def save(name, **d):
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **npy(d))
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB")


class _Vocab:                       # beam_search_predict only needs len() and .stoi
    stoi = {"<PAD>": 0, "<SOS>": 1, "<EOS>": 2, "<UNK>": 3}

    def __init__(self, n):
        self.n = n

    def __len__(self):
        return self.n


def gen_pickles():
    torch.manual_seed(21)
    B, T, L, V, F, H, E, A = 5, 6, 7, 23, 16, 24, 12, 8
    dec = FeaturesCaptioning(in_feature_size=F, output_size=V, rnn_hidden_size=H, embedding_size=E, attn_size=A,
                             rnn_dropout=0.0)
    with torch.no_grad():
        dec.out.weight.mul_(6.0)
    grec = GlobalReconstructor(decoder_size=H, hidden_size=F)
    lrec = LocalReconstructor(decoder_size=H, hidden_size=F, attn_size=A)
    torch.save({"decoder": dec, "global": grec, "local": lrec}, os.path.join(OUT, "ref_decoder_tiny.pt"))
    print("ref_decoder_tiny.pt:", os.path.getsize(os.path.join(OUT, "ref_decoder_tiny.pt")) / 1024, "KiB")
    g = torch.Generator().manual_seed(22)
    feats = torch.relu(torch.randn(B, T, F, generator=g)) * 2.0
    _, _, caps = O.synth_batch(B, T, L, V, Fa=1, Fv=1, seed=23, min_cap=3)
    with torch.no_grad():
        out, hid = dec.decode(feats, caps, L, 1.0)
        free, _ = dec.decode(feats, None, 9)
        beam = dec.beam_search_predict(feats, _Vocab(V), max_caption_len=8, beam_alpha=0, beam_width=3)
        beam = torch.tensor([[int(x) for x in row] for row in beam])
        g_rec = grec.reconstruct(hid, out, caps, T)
        l_rec = lrec.reconstruct(hid, out, caps, T)
    save("ref_pickle_expect", feats=feats, caps=caps, out=out, hid=hid, greedy_ids=free.argmax(2).t(), beam_ids=beam,
         g_rec=g_rec, l_rec=l_rec)


def gen_word_step():
    torch.manual_seed(31)
    B, T, V, F, H, E, A, L = 4, 5, 19, 12, 16, 10, 8, 6
    dec = FeaturesCaptioning(in_feature_size=F, output_size=V, rnn_hidden_size=H, embedding_size=E, attn_size=A,
                             rnn_dropout=0.0)
    g = torch.Generator().manual_seed(32)
    feats = (torch.relu(torch.randn(B, T, F, generator=g)) * 1.5).requires_grad_()
    h0 = (torch.randn(1, B, H, generator=g) * 0.5).requires_grad_()
    c0 = (torch.randn(1, B, H, generator=g) * 0.5).requires_grad_()
    words = torch.randint(0, V, (1, B), generator=g)
    wlogp = torch.randn(B, V, generator=g)
    wh, wc = torch.randn(1, B, H, generator=g), torch.randn(1, B, H, generator=g)
    logp, (h1, c1), alpha = dec.forward_word(feats, (h0, c0), words)
    ((logp * wlogp).sum() + (h1 * wh).sum() + (c1 * wc).sum()).backward()
    d = {"feats": feats, "h0": h0, "c0": c0, "words": words, "wlogp": wlogp, "wh": wh, "wc": wc, "logp": logp,
         "h1": h1, "c1": c1, "alpha": alpha, "dfeats": feats.grad, "dh0": h0.grad, "dc0": c0.grad}
    d.update({"p." + k: v.detach().clone() for k, v in dec.state_dict().items()})
    d.update({"g." + k: v.grad.clone() for k, v in dec.named_parameters()})
    # forward_sentence from a non-zero state (scheduled sampling draws from the global CPU RNG)
    dec.zero_grad()
    _, _, caps = O.synth_batch(B, T, L, V, Fa=1, Fv=1, seed=33, min_cap=3)
    wsent = torch.randn(L, B, V, generator=g)
    torch.manual_seed(34)
    sent, hids = dec.forward_sentence(feats.detach(), caps, (h0.detach(), c0.detach()), L, 0.5)
    (sent * wsent).sum().backward()
    d.update({"caps": caps, "wsent": wsent, "sent": sent, "hids": hids, "sent_rng_after": torch.rand(1)})
    d.update({"gs." + k: v.grad.clone() for k, v in dec.named_parameters()})
    # stand-alone masked attention with gradients
    att = TemporalAttention(hidden_size=H, feature_size=F, bottleneck_size=A)
    q = (torch.randn(B, H, generator=g)).requires_grad_()
    keys = (torch.randn(B, T, F, generator=g)).requires_grad_()
    mask = torch.ones(B, T, dtype=torch.bool)
    mask[0, 3:] = False
    mask[2, 1:] = False
    wctx = torch.randn(B, F, generator=g)
    ctx, w = att(q, keys, mask)
    (ctx * wctx).sum().backward()
    d.update({"att.q": q, "att.keys": keys, "att.mask": mask, "att.wctx": wctx, "att.ctx": ctx, "att.alpha": w,
              "att.dq": q.grad, "att.dkeys": keys.grad})
    d.update({"att.p." + k: v.detach().clone() for k, v in att.state_dict().items()})
    d.update({"att.g." + k: v.grad.clone() for k, v in att.named_parameters()})
    save("word_step_small", **d)


def gen_total_loss():
    g = torch.Generator().manual_seed(41)
    B, T, L, V, F = 4, 5, 6, 17, 9
    _, _, caps = O.synth_batch(B, T, L, V, Fa=1, Fv=1, seed=42, min_cap=3)
    d = {"caps": caps}
    logits = torch.randn(L, B, V, generator=g)
    feats = torch.relu(torch.randn(B, T, F, generator=g))
    d["feats"] = feats
    for kind in ("none", "global", "local"):
        out = torch.log_softmax(logits, 2).clone().requires_grad_()
        rec = None
        if kind != "none":
            rec = torch.randn(B, L if kind == "global" else T, F, generator=g).requires_grad_()
        fn = ref_losses.ReconstructionLossBuilder(reg_lambda=0.0005, recon_lambda=0.5, reconstruction_type=kind)
        terms = fn(out, caps, feats, rec)
        terms[0].mean().backward()
        d[f"{kind}.out"] = out
        d[f"{kind}.terms"] = torch.stack([t.reshape(()) if t.numel() == 1 else t for t in terms])
        d[f"{kind}.loss_shape"] = np.asarray(terms[0].shape, dtype=np.int64)
        d[f"{kind}.dout"] = out.grad
        if rec is not None:
            d[f"{kind}.rec"] = rec
            d[f"{kind}.drec"] = rec.grad
    save("total_loss_small", **d)


def gen_nlp_scores():
    """Seeded sentence sets scored by the reference's own pycocoevalcap scorers (Bleu / Rouge / Cider, the three that
    need no Java) -> tests/golden/nlp_scores_small.json, the pin of salstm/nlp_score.py."""
    import contextlib
    import io
    import json
    import random
    from pycocoevalcap.bleu.bleu import Bleu
    from pycocoevalcap.cider.cider import Cider
    from pycocoevalcap.rouge.rouge import Rouge
    rng = random.Random(0)
    words = [f"w{i}" for i in range(30)]

    def sent(lo, hi):
        return " ".join(rng.choice(words) for _ in range(rng.randint(lo, hi)))
    cases = []
    for _ in range(4):
        gts, res = {}, {}
        for v in range(rng.randint(3, 12)):
            refs = [sent(3, 12) for _ in range(rng.randint(1, 5))]
            base = refs[0].split()
            hyp = " ".join(w if rng.random() < 0.7 else rng.choice(words) for w in base[:rng.randint(1, len(base))])
            gts[f"vid{v}"] = refs
            res[f"vid{v}"] = [hyp]
        cases.append((gts, res))
    cases.append(({"a": ["w1 w2 w3"]}, {"a": ["w9"]}))
    out = []
    for gts, res in cases:
        with contextlib.redirect_stdout(io.StringIO()):
            rb, _ = Bleu(4).compute_score(gts, res)
            rr, _ = Rouge().compute_score(gts, res)
            rc, _ = Cider().compute_score(gts, res)
        out.append({"gts": gts, "res": res, "Bleu": list(rb), "ROUGE_L": float(rr), "CIDEr": float(rc)})
    json.dump(out, open(os.path.join(OUT, "nlp_scores_small.json"), "w"))
    print("nlp_scores_small.json:", len(out), "cases")


if __name__ == "__main__":
    gen_nlp_scores()
    gen_pickles()
    gen_word_step()
    gen_total_loss()
