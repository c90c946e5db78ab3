"""GPU tests of the drop-in surface: whole-module pickles, word-by-word stepping with gradients,
the single-stream loss twin, backward through individual loss terms, the gradient arena, and the reference's own
`Trainer` (src/train.py, staged under oracle/_ref by oracle/build_ref.py) driving the B200 modules unchanged."""
import os
import sys
import types

import pytest
import torch

from conftest import GOLDEN, PKG, ROOT, load_golden, sub
from oracle import salstm_oracle as O

pytestmark = pytest.mark.gpu


def close(a, b, atol=1e-4, rtol=1e-4):
    torch.testing.assert_close(a.detach().cpu().float(), b.detach().cpu().float(), atol=atol, rtol=rtol)


@pytest.fixture(scope="module")
def dev():
    import __graft_entry__ as g
    g.build()
    from salstm import cabi
    assert cabi.lib().mvc_device_ok() == 1, "not an sm_100 device"
    return torch.device("cuda:0")


class Vocab:
    def __init__(self, n):
        self.itos = {0: "<PAD>", 1: "<SOS>", 2: "<EOS>", 3: "<UNK>"}
        for i in range(4, n):
            self.itos[i] = f"w{i}"
        self.stoi = {v: k for k, v in self.itos.items()}

    def __len__(self):
        return len(self.itos)

    def decode_indexes(self, idx):
        return O.decode_indexes(self.itos, idx)


def _prefix(ids):
    out = []
    for row in ids:
        row = [int(x) for x in row]
        out.append(row[:row.index(2) + 1] if 2 in row[1:] else row)
    return out


# --------------------------------------------------------------------------- pickles (train.py:162-173, notebook)
def test_reference_pickle_predicts_on_gpu(dev):
    """A whole-module pickle written by the unmodified reference loads into the repo classes, moves to the GPU and
    reproduces the reference's own outputs (fp32 exact path by default)."""
    import models  # noqa: F401
    blob = torch.load(os.path.join(GOLDEN, "ref_decoder_tiny.pt"), weights_only=False)
    g = load_golden("ref_pickle_expect")
    dec, grec, lrec = (blob[k].to(dev) for k in ("decoder", "global", "local"))
    feats, caps = g["feats"].to(dev), g["caps"].to(dev)
    with torch.no_grad():
        out, hid = dec.decode(feats, caps, caps.shape[0], 1.0)
        close(out, g["out"]); close(hid, g["hid"])
        assert torch.equal(dec.greedy_ids(feats, 9).cpu(), g["greedy_ids"])
        beam = dec.beam_search_predict(feats, Vocab(23), max_caption_len=8, beam_alpha=0, beam_width=3)
        assert _prefix(beam) == _prefix(g["beam_ids"].tolist())
        close(grec.reconstruct(hid, out, caps, feats.shape[1]), g["g_rec"])
        close(lrec.reconstruct(hid, out, caps, feats.shape[1]), g["l_rec"])


def test_full_model_pickle_from_staged_reference(dev, tmp_path):
    """torch.save(<reference AVCaptioningDual>) -> torch.load with the repo's `models` package -> predict()."""
    from oracle import build_ref
    if not build_ref.available():
        pytest.skip("oracle/_ref not staged")
    ref = build_ref.load()
    import models
    V = 37
    torch.manual_seed(3)
    with build_ref._Saved():                           # the reference's classes must be importable while pickling
        sys.modules["models"] = types.ModuleType("models"); sys.modules["models"].__path__ = [os.path.join(ref.root, "models")]
        import importlib
        for sub_ in ("temporal_attention", "features_captioning", "reconstructor", "captioning"):
            importlib.import_module("models." + sub_)
        cap = sys.modules["models.captioning"]
        rmodel = cap.AVCaptioningDual(Vocab(V), 1.0, "global", device="cpu")
        path = tmp_path / "ref_model_best.pt"
        torch.save(rmodel, path)                                                     # train.py:162
        a, v, caps = O.synth_batch(3, 5, 6, V, seed=4, min_frames=2, min_cap=3)
        with torch.no_grad():
            r_out, r_ar, r_vr = rmodel(a, v, caps, teacher_forcing_ratio=1.0)
            r_txt = rmodel.predict(a, v, max_caption_len=7, mode="direct")
    model = torch.load(path, weights_only=False)                                     # notebook predict_captions
    assert type(model) is models.AVCaptioningDual and type(model.v_decoder) is models.FeaturesCaptioning
    model = model.to(dev)
    with torch.no_grad():
        out, ar, vr = model(a.to(dev), v.to(dev), caps.to(dev), teacher_forcing_ratio=1.0)
        close(out, r_out, atol=2e-4); close(ar, r_ar, atol=2e-4); close(vr, r_vr, atol=2e-4)
        assert model.predict(a.to(dev), v.to(dev), max_caption_len=7, mode="direct") == r_txt
    # and back: a pickle written from here names models.captioning.*, not the implementation module
    path2 = tmp_path / "b200_model_last.pt"
    torch.save(model, path2)
    assert b"salstm" not in open(path2, "rb").read()
    again = torch.load(path2, weights_only=False)
    assert type(again) is models.AVCaptioningDual


# --------------------------------------------------------------------------- word-by-word stepping (features_captioning.py:77-119)
def _tiny_decoder(dev, g):
    from models import FeaturesCaptioning
    p = sub(g, "p.")
    dec = FeaturesCaptioning(in_feature_size=p["attention.U.weight"].shape[1], output_size=p["out.weight"].shape[0],
                             rnn_hidden_size=p["out.weight"].shape[1], embedding_size=p["embedding.weight"].shape[1],
                             attn_size=p["attention.b"].shape[0], device=dev).to(dev)
    dec.load_state_dict(p)
    return dec


def test_forward_word_is_differentiable(dev):
    g = load_golden("word_step_small")
    dec = _tiny_decoder(dev, g)
    feats = g["feats"].to(dev).requires_grad_()
    h0, c0 = g["h0"].to(dev).requires_grad_(), g["c0"].to(dev).requires_grad_()
    logp, (h1, c1), alpha = dec.forward_word(feats, (h0, c0), g["words"].to(dev))
    close(logp, g["logp"]); close(h1, g["h1"]); close(c1, g["c1"]); close(alpha, g["alpha"])
    ((logp * g["wlogp"].to(dev)).sum() + (h1 * g["wh"].to(dev)).sum() + (c1 * g["wc"].to(dev)).sum()).backward()
    close(feats.grad, g["dfeats"], rtol=1e-3); close(h0.grad, g["dh0"], rtol=1e-3); close(c0.grad, g["dc0"], rtol=1e-3)
    for k, v in dec.named_parameters():
        close(v.grad, g["g." + k], atol=2e-5, rtol=1e-3)


def test_forward_sentence_from_nonzero_state(dev):
    """A non-zero initial state steps word by word (same RNG draws, differentiable); the zero state of _init_hidden
    takes the fused kernels -- both against the reference's outputs."""
    g = load_golden("word_step_small")
    dec = _tiny_decoder(dev, g)
    caps = g["caps"].to(dev)
    L = caps.shape[0]
    torch.manual_seed(34)
    sent, hids = dec.forward_sentence(g["feats"].to(dev), caps, (g["h0"].to(dev), g["c0"].to(dev)), L, 0.5)
    close(torch.rand(1), g["sent_rng_after"], atol=0, rtol=0)                 # consumed exactly L-1 draws
    close(sent, g["sent"]); close(hids, g["hids"])
    (sent * g["wsent"].to(dev)).sum().backward()
    for k, v in dec.named_parameters():
        close(v.grad, g["gs." + k], atol=5e-5, rtol=2e-3)
    with torch.no_grad():
        a, b = dec.forward_sentence(g["feats"].to(dev), caps, dec._init_hidden(caps.shape[1]), L, 1.0)
        c, d = dec.decode(g["feats"].to(dev), caps, L, 1.0)
    assert torch.equal(a, c) and torch.equal(b, d)


def test_standalone_attention_is_differentiable(dev):
    from models import TemporalAttention
    g = load_golden("word_step_small")
    p = sub(g, "att.p.")
    att = TemporalAttention(p["W.weight"].shape[1], p["U.weight"].shape[1], p["b"].shape[0]).to(dev)
    att.load_state_dict(p)
    q, keys = g["att.q"].to(dev).requires_grad_(), g["att.keys"].to(dev).requires_grad_()
    ctx, w = att(q, keys, g["att.mask"].to(dev))
    close(ctx, g["att.ctx"]); close(w, g["att.alpha"])
    (ctx * g["att.wctx"].to(dev)).sum().backward()
    close(q.grad, g["att.dq"], rtol=1e-3); close(keys.grad, g["att.dkeys"], rtol=1e-3)
    for k, v in att.named_parameters():
        close(v.grad, g["att.g." + k], atol=2e-5, rtol=1e-3)


# --------------------------------------------------------------------------- losses
@pytest.mark.parametrize("kind", ["none", "global", "local"])
def test_total_reconstruction_loss_twin(dev, kind):
    """losses.py:43-83 (single-stream twin of the modality-wise loss): values, shapes and gradients."""
    import losses as L
    g = load_golden("total_loss_small")
    out = g[f"{kind}.out"].to(dev).requires_grad_()
    rec = g[f"{kind}.rec"].to(dev).requires_grad_() if kind != "none" else None
    fn = L.ReconstructionLossBuilder(reg_lambda=0.0005, recon_lambda=0.5, reconstruction_type=kind)
    terms = fn(out, g["caps"].to(dev), g["feats"].to(dev), rec)
    assert len(terms) == 4 and tuple(terms[0].shape) == tuple(int(x) for x in g[f"{kind}.loss_shape"])
    close(torch.stack([t.reshape(()) for t in terms]), g[f"{kind}.terms"], atol=1e-6, rtol=2e-5)
    terms[0].mean().backward()
    close(out.grad, g[f"{kind}.dout"], atol=1e-7, rtol=1e-4)
    if rec is not None:
        close(rec.grad, g[f"{kind}.drec"], atol=1e-8, rtol=1e-4)


def test_backward_through_individual_loss_terms(dev):
    """The reference returns graph tensors for ce / entropy / reconstruction terms; (3*ce + 2*loss - ent + v_rec).backward()
    must give the oracle's gradients, and upstream gradients other than 1 are honoured."""
    import losses as L
    g = torch.Generator().manual_seed(5)
    B, T, Lc, V, F = 4, 5, 6, 17, 12
    _, _, caps = O.synth_batch(B, T, Lc, V, Fa=1, Fv=1, seed=6, min_cap=3)
    logits = torch.randn(Lc, B, V, generator=g)
    feats = torch.relu(torch.randn(B, T, F, generator=g))
    rec0 = torch.randn(B, T, F, generator=g)
    lam = dict(reg_lambda=0.0005, audio_recon_lambda=0.25, visual_recon_lambda=0.5)

    def run(device, lossmod):
        out = torch.log_softmax(logits, 2).to(device).requires_grad_()
        rec = rec0.to(device).requires_grad_()
        f = feats.to(device)
        t = lossmod(out, caps.to(device), f[..., :4], rec[..., :4], f[..., 4:], rec[..., 4:], rec_type="local", **lam)
        (3.0 * t[1] + 2.0 * t[0] - t[2] + t[4]).backward()
        return out.grad, rec.grad, torch.stack([x.detach() for x in t])

    go, gr, tv = run(dev, L.ModalityWiseReconstructionLoss)
    oo, orr, otv = run("cpu", O.modality_wise_loss)
    close(tv, otv, atol=1e-6, rtol=2e-5)
    close(go, oo, atol=1e-7, rtol=1e-4); close(gr, orr, atol=1e-8, rtol=1e-4)
    # plain path with an upstream gradient != 1 (scaled in place by the device scalar), and a refused second backward
    out = torch.log_softmax(logits, 2).to(dev).requires_grad_()
    t = L.ModalityWiseReconstructionLoss(out, caps.to(dev), reg_lambda=0.0005)
    (t[0] * 0.125).backward(retain_graph=True)
    ocpu = torch.log_softmax(logits, 2).requires_grad_()
    (O.modality_wise_loss(ocpu, caps, reg_lambda=0.0005)[0] * 0.125).backward()
    close(out.grad, ocpu.grad, atol=1e-8, rtol=1e-4)
    with pytest.raises(RuntimeError, match="second backward"):
        t[0].backward()


# --------------------------------------------------------------------------- gradient arena / optimizer
def _small_model(dev, precision="fp32"):
    from models import FeaturesCaptioning
    torch.manual_seed(11)
    return FeaturesCaptioning(in_feature_size=24, output_size=41, rnn_hidden_size=32, embedding_size=16, attn_size=16,
                              device=dev, precision=precision).to(dev)


def _loss(dec, feats, caps):
    out, _ = dec.decode(feats, caps, caps.shape[0], 1.0)
    return torch.nn.functional.nll_loss(out[1:].reshape(-1, out.shape[2]), caps[1:].reshape(-1), ignore_index=0)


def test_flat_clip_adam_multi_step_matches_torch_adam(dev):
    """FlatClipAdam + gradient arena over several steps == clip_grad_value_ + torch.optim.Adam(amsgrad) on a twin
    (train.py:86-88, 207-210), including a parameter that never gets a gradient and lr decay by a scheduler."""
    from salstm.trainer import FlatClipAdam
    a, b = _small_model(dev), _small_model(dev)
    b.load_state_dict(a.state_dict())
    unused_a, unused_b = (torch.nn.Parameter(torch.ones(5, device=dev)) for _ in range(2))
    oa = FlatClipAdam(list(a.parameters()) + [unused_a], lr=1e-2, weight_decay=1e-5, clip_value=0.05)
    ob = torch.optim.Adam(list(b.parameters()) + [unused_b], lr=1e-2, weight_decay=1e-5, amsgrad=True)
    sa = torch.optim.lr_scheduler.ReduceLROnPlateau(oa, factor=0.5, patience=0)
    sb = torch.optim.lr_scheduler.ReduceLROnPlateau(ob, factor=0.5, patience=0)
    for step in range(5):
        feats, _, caps = O.synth_batch(6, 5, 7, 41, Fa=24, Fv=1, seed=20 + step, min_frames=2, min_cap=3)
        feats, caps = (feats / 255.0).to(dev), caps.to(dev)
        oa.zero_grad(); ob.zero_grad()
        la, lb = _loss(a, feats, caps), _loss(b, feats, caps)
        la.backward(); lb.backward()
        if step >= 1:                       # after the first step the kernels write into the arena views
            assert a.out.weight.grad.data_ptr() == oa._views[[id(p) for p in oa._live].index(id(a.out.weight))].data_ptr()
        torch.nn.utils.clip_grad_value_(b.parameters(), 0.05)
        oa.step(); ob.step()
        sa.step(float(step)); sb.step(float(step))      # metric gets worse every step -> lr halves
        for (k, pa), pb in zip(a.named_parameters(), b.parameters()):
            close(pa, pb, atol=2e-6, rtol=1e-5)
    assert oa.lr == pytest.approx(ob.param_groups[0]["lr"]) and oa.lr < 1e-2
    assert torch.equal(unused_a, unused_b) and unused_a.grad is None
    sd = oa.state_dict()
    assert sd["step"] == 5 and sd["exp_avg"].numel() == sum(p.numel() for p in a.parameters())


def test_arena_accumulation_semantics(dev):
    """Two backwards without zero_grad, and zero_grad(set_to_none=False): gradients ACCUMULATE exactly like plain
    autograd tensors even though the arena is registered (no self-aliasing `p.grad += p.grad`)."""
    from salstm.trainer import FlatClipAdam
    a, b = _small_model(dev), _small_model(dev)
    b.load_state_dict(a.state_dict())
    opt = FlatClipAdam(a.parameters(), lr=0.0, weight_decay=0.0)
    feats, _, caps = O.synth_batch(6, 5, 7, 41, Fa=24, Fv=1, seed=30, min_frames=2, min_cap=3)
    feats, caps = (feats / 255.0).to(dev), caps.to(dev)
    _loss(a, feats, caps).backward(); opt.step(); opt.zero_grad()            # flatten + register the arena
    _loss(a, feats, caps).backward(); _loss(a, feats, caps).backward()      # accumulate twice
    _loss(b, feats, caps).backward()
    for pa, pb in zip(a.parameters(), b.parameters()):
        close(pa.grad, 2.0 * pb.grad, atol=1e-6, rtol=1e-5)
    opt.zero_grad(set_to_none=False)
    assert float(opt.flat_g.abs().sum()) == 0.0
    _loss(a, feats, caps).backward()
    for pa, pb in zip(a.parameters(), b.parameters()):
        close(pa.grad, pb.grad, atol=1e-6, rtol=1e-5)
        assert pa.grad.data_ptr() >= opt.flat_g.data_ptr()                  # still the flat buffer's views


# --------------------------------------------------------------------------- the reference's own Trainer (src/train.py)
def _import_reference_train():
    """`import train` from the staged reference with stub spacy / tensorboardX (absent here) and the repo's
    package AHEAD of the reference's src/, exactly what salstm/launch.py arranges for `python src/train.py`."""
    from oracle import build_ref
    if not build_ref.available() or not os.path.isfile(os.path.join(build_ref.OUT, "src", "train.py")):
        pytest.skip("oracle/_ref (train.py, get_loader.py) not staged")
    spacy = types.ModuleType("spacy")

    class _Tok:
        def __init__(self, t):
            self.text = t

    class _NLP:
        def tokenizer(self, text):
            return [_Tok(t) for t in text.split()]
    spacy.load = lambda name: _NLP()
    tbx = types.ModuleType("tensorboardX")

    class SummaryWriter:
        def __init__(self, *a, **k):
            self.scalars = []

        def add_scalar(self, tag, value, step):
            self.scalars.append((tag, float(value), int(step)))
    tbx.SummaryWriter = SummaryWriter
    sys.modules.setdefault("spacy", spacy)
    sys.modules.setdefault("tensorboardX", tbx)
    src = os.path.join(build_ref.OUT, "src")
    sys.path[:] = [p for p in sys.path if p != src] + [src]          # behind PKG: models/losses resolve to the repo's
    assert sys.path.index(PKG) < sys.path.index(src)
    for name in ("train", "get_loader"):
        sys.modules.pop(name, None)
    import train
    import models
    import losses
    assert os.path.dirname(models.__file__) == os.path.join(PKG, "models") and os.path.dirname(losses.__file__) == PKG
    assert train.AVCaptioningDual is models.AVCaptioningDual
    return train


@pytest.mark.parametrize("rec_type", ["none", "global"])
def test_reference_trainer_runs_unchanged_over_the_b200_modules(dev, tmp_path, rec_type):
    """src/train.py's Trainer.train (2 optimiser steps: zero_grad, forward, RecLoss, backward, clip_grad_value_,
    Adam(amsgrad).step, the .item() logging), Trainer.test (tf_ratio=0) and Trainer.eval (predict -> strings), all
    unmodified, over AVCaptioningDual (`dual = True`, train.py:375) -- and the loss curve equals the oracle's."""
    train = _import_reference_train()
    import get_loader
    V = 53
    vocab = get_loader.Vocabulary(freq_threshold=1)
    for i in range(4, V):
        vocab.itos[i] = f"w{i}"; vocab.stoi[f"w{i}"] = i
    torch.manual_seed(0)
    model = train.AVCaptioningDual(vocab=vocab, teacher_forcing_ratio=1.0, reconstructor_type=rec_type, device=dev)
    model.to(dev)
    p0 = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    batches = [O.synth_batch(4, 5, 6, V, seed=50 + i, min_frames=2, min_cap=3) for i in range(2)]
    batches = [(a / 255.0, v / 10.0, c) for a, v, c in batches]
    tr = train.Trainer(checkpoint_name=str(tmp_path / "ck" / "m.ckpt"), log_dir=str(tmp_path / "logs"), display_freq=1)
    cfg = train.TrainerConfig()
    tr.device = dev
    tr.optimizer = cfg.optimizer(model.parameters(), lr=cfg.lr, weight_decay=cfg.weight_decay, amsgrad=True)   # :86-88
    tr.gradient_clip_value = cfg.gradient_clip_value
    lam = dict(reg_lambda=0.0005, audio_recon_lambda=0.00005, visual_recon_lambda=0.5)
    tr.RecLoss = train.ModalityWiseReconstructionLossBuilder(rec_type=model.reconstructor_type, **lam)          # :103-108
    tr.history = {}
    hist = tr.train(model, batches, epoch=1)
    val = tr.test(model, batches, "val", epoch=1)
    assert all(torch.isfinite(torch.tensor(v)) for v in list(hist.values()) + list(val.values()))
    # the same two steps with the oracle on the CPU: identical loss trajectory
    p = {k: v.clone().requires_grad_() for k, v in p0.items() if not k.startswith("output_fc")}
    opt = torch.optim.Adam(list(p.values()), lr=cfg.lr, weight_decay=cfg.weight_decay, amsgrad=True)
    tot = 0.0
    for a, v, c in batches:
        opt.zero_grad()
        out, ar, vr = O.av_dual_forward(p, a, v, c, 1.0, rec_type)
        t = O.modality_wise_loss(out, c, a, ar, v, vr, rec_type=rec_type, **lam)
        t[0].backward()
        torch.nn.utils.clip_grad_value_(list(p.values()), cfg.gradient_clip_value)
        opt.step()
        tot += float(t[0])
    assert hist["total"] == pytest.approx(tot / 2, rel=1e-4)
    tags = {t for t, _, _ in tr.summary_writer.scalars}
    assert {"train/loss", "train/loss/ce", "val/loss", "train_epoch/loss/v_recon"} <= tags
    tr._save_checkpoint(1, model, {})                                                                          # :66-84
    assert sorted(torch.load(tmp_path / "ck" / "m.ckpt")["v_decoder"]) == sorted(model.v_decoder.state_dict())
    # Trainer.eval: predict() -> strings per video id
    loader = [(["vid0", "vid1", "vid2", "vid3"], batches[0][0], batches[0][1], [["w5 w6 w7", "w5 w9"]] * 4)]
    scores, gt, gen = tr.eval(model, loader, "val", 1, mode="direct", get_scores=True)      # NLPScore: train.py:337-346
    assert set(gen) == {"vid0", "vid1", "vid2", "vid3"} and all(isinstance(v[0], str) for v in gen.values())
    assert {"Bleu_1", "Bleu_4", "ROUGE_L", "CIDEr", "METEOR"} <= set(scores) and 0.0 <= scores["Bleu_1"] <= 1.0
    assert "val/score/direct/CIDEr" in {t for t, _, _ in tr.summary_writer.scalars}
    with torch.no_grad():
        ids = O.av_dual_greedy_ids({k: v.detach().cpu() for k, v in model.state_dict().items()}, batches[0][0],
                                   batches[0][1], 30)
    assert [gen[f"vid{i}"][0] for i in range(4)] == [vocab.decode_indexes(r[1:]) for r in ids.tolist()]


def test_graphed_train_step_equals_eager_steps(dev):
    """salstm.trainer.GraphedTrainStep (one optimiser step as a CUDA-graph replay, device-side step count / lr) gives
    the parameters the eager loop gives, step after step, on the bf16 persistent-kernel path, including an lr change."""
    from models import AVCaptioning
    from salstm.trainer import FlatClipAdam, GraphedTrainStep
    import losses as L
    V, B, T, Lc = 97, 8, 6, 7
    lam = dict(reg_lambda=0.0005, audio_recon_lambda=0.0, visual_recon_lambda=0.0)

    def make():
        torch.manual_seed(5)
        m = AVCaptioning(Vocab(V), 1.0, "none", device=dev, precision="bf16").to(dev)
        return m, FlatClipAdam(m.parameters(), lr=1e-3, weight_decay=1e-5, clip_value=5.0)

    batches = []
    for i in range(9):
        a, v, c = O.synth_batch(B, T, Lc, V, seed=60 + i, min_frames=2, min_cap=3)
        batches.append(((a / 255.0).to(dev), (v / 10.0).to(dev), c.to(dev)))
    loss_fn = L.ModalityWiseReconstructionLossBuilder(rec_type="none", **lam)
    ma, oa = make()
    mb, ob = make()
    gstep = GraphedTrainStep(ma, loss_fn, oa, batches[0], warmup=2)       # runs 2 eager warm-up steps on batches[0]
    losses_a, losses_b = [], []
    for _ in range(2):                                                    # the twin takes the same two warm-up steps
        ob.zero_grad()
        out, ar, vr = mb(*batches[0])
        loss_fn(out, batches[0][2], batches[0][0], ar, batches[0][1], vr)[0].mean().backward()
        ob.step()
    for i, (a, v, c) in enumerate(batches[1:7]):
        if i == 3:
            oa.param_groups[0]["lr"] = 2.5e-4; ob.param_groups[0]["lr"] = 2.5e-4     # ReduceLROnPlateau-style change
        losses_a.append(float(gstep(a, v, c)[0]))
        ob.zero_grad()
        out, ar, vr = mb(a, v, c)
        t = loss_fn(out, c, a, ar, v, vr)
        t[0].mean().backward()
        ob.step()
        losses_b.append(float(t[0]))
    assert losses_a == pytest.approx(losses_b, rel=1e-6)
    assert oa.step_count == ob.step_count == 8
    for (k, pa), pb in zip(ma.named_parameters(), mb.parameters()):
        # same kernels in the same order; only the embedding scatter (atomicAdd order) and the bias-correction powf
        # (device vs host libm) may differ in the last bits
        torch.testing.assert_close(pa, pb, rtol=2e-5, atol=2e-6, msg=k)


def test_graphed_train_step_input_slots_in_place(dev, tmp_path):
    """GraphedTrainStep(slots=2): batches written straight into the graphs' static input tensors -- here by a ShardFeeder
    with device_slots= -- take the copy-free replay and give the eager loop's parameters."""
    from models import AVCaptioning
    from salstm.trainer import FlatClipAdam, GraphedTrainStep
    from salstm.shards import ShardFeeder, ShardReader, write_shard
    import losses as L
    V, B, T, Lc, N = 97, 8, 6, 7, 40
    lam = dict(reg_lambda=0.0005, audio_recon_lambda=0.0, visual_recon_lambda=0.0)
    a, v, c = O.synth_batch(N, T, Lc, V, seed=77, min_frames=2, min_cap=3)
    a, v = a / 255.0, v / 10.0
    path = str(tmp_path / "s.shard")
    write_shard(path, list(a), list(v), list(c.t()), T=T, L=Lc)

    def make():
        torch.manual_seed(6)
        m = AVCaptioning(Vocab(V), 1.0, "none", device=dev, precision="bf16").to(dev)
        return m, FlatClipAdam(m.parameters(), lr=1e-3, weight_decay=1e-5, clip_value=5.0)

    loss_fn = L.ModalityWiseReconstructionLossBuilder(rec_type="none", **lam)
    first = (a[:B].bfloat16().to(dev), v[:B].bfloat16().to(dev), c[:, :B].contiguous().to(dev))
    ma, oa = make()
    mb, ob = make()
    gstep = GraphedTrainStep(ma, loss_fn, oa, first, warmup=2, slots=2)
    assert len(gstep.input_slots) == 2
    for _ in range(2):
        ob.zero_grad()
        out, ar, vr = mb(*first)
        loss_fn(out, first[2], first[0], ar, first[1], vr)[0].mean().backward()
        ob.step()
    feeder = ShardFeeder(ShardReader(path, pin=True), B, dev, shuffle=False, device_slots=gstep.input_slots)
    la, lb, n = [], [], 0
    for fa, fv, fc, _ in feeder:
        k = n % 2
        assert fa.data_ptr() == gstep.input_slots[k][0].data_ptr()          # uploaded in place
        ref = tuple(t.clone() for t in (fa, fv, fc))
        la.append(float(gstep(fa, fv, fc)[0]))
        ob.zero_grad()
        out, ar, vr = mb(*ref)
        t = loss_fn(out, ref[2], ref[0], ar, ref[1], vr)
        t[0].mean().backward()
        ob.step()
        lb.append(float(t[0]))
        n += 1
    assert n == N // B
    assert la == pytest.approx(lb, rel=1e-6)
    assert oa.step_count == ob.step_count == 2 + n
    for (k, pa), pb in zip(ma.named_parameters(), mb.parameters()):
        torch.testing.assert_close(pa, pb, rtol=2e-5, atol=2e-6, msg=k)
    with pytest.raises(ValueError):
        f32 = (a[:B].to(dev), v[:B].to(dev), c[:, :B].contiguous().to(dev))
        ShardFeeder(ShardReader(path, pin=True), B, dev, device_slots=[f32, f32])          # slots must be bf16 features
