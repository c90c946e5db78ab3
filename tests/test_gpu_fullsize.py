"""Parity at the BENCHMARKED shapes: the exact kernels / instantiations bench.py times, against
the CPU oracle on the same seeded inputs.

  * C2 (BASELINE configs[1]): AVCaptioning joint, bf16, B=128 T=44 L=24 V=3201 F=2176 -- the persistent recurrence
    kernels `recur_*_kernel<8,9>` + tcgen05 GEMMs + fused loss, vs the oracle in fp64: log-probs atol 5e-2, loss rtol
    2e-2, per-parameter gradient cosine >= 0.999 and norms within 5 % (unit-scaled features).
  * C3: fp32 exact path greedy ids at B=512 T=30 L=30 V=10547 (peaky logits) identical to the oracle's.
  * The same C2 step on the fp32 exact path vs the fp32 oracle at rtol 1e-3 on gradients.
"""
import pytest
import torch

from oracle import salstm_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    import __graft_entry__ as g
    g.build()
    from salstm import cabi
    assert cabi.lib().mvc_device_ok() == 1, "not an sm_100 device"
    return torch.device("cuda:0")


class Vocab:
    stoi = {"<PAD>": 0, "<SOS>": 1, "<EOS>": 2, "<UNK>": 3}

    def __init__(self, n):
        self.n = n

    def __len__(self):
        return self.n


def cos(a, b):
    a, b = a.detach().cpu().double().flatten(), b.detach().cpu().double().flatten()
    return float((a @ b) / (a.norm() * b.norm() + 1e-30))


def _load(model, params):
    sd = model.state_dict()
    sd.update({k: params[k].clone() for k in sd if k in params})
    model.load_state_dict(sd)


LAM = dict(reg_lambda=0.0005, audio_recon_lambda=0.00005, visual_recon_lambda=0.5)     # train.py:412-461


@pytest.fixture(scope="module")
def c2_oracle():
    """fp64 oracle forward + loss + backward at the C2 shape (a few tens of seconds of CPU)."""
    B, T, L, V = 128, 44, 24, 3201
    gen = torch.Generator().manual_seed(0)
    p = O.init_decoder_params("decoder.", 2176, V, gen=gen)
    audio, visual, caps = O.synth_batch(B, T, L, V, seed=1)
    audio, visual = audio / 255.0, visual / 10.0
    pd = {k: v.double().requires_grad_() for k, v in p.items()}
    out, _, _ = O.av_forward(pd, audio.double(), visual.double(), caps, 1.0, "none", hoist=True)
    terms = O.modality_wise_loss(out, caps, **LAM)
    terms[0].backward()
    return p, (audio, visual, caps), out.detach(), [float(t) for t in terms], {k: v.grad for k, v in pd.items()}


def test_c2_bf16_persistent_kernels_vs_fp64_oracle(dev, c2_oracle):
    from models import AVCaptioning
    import losses as Lm
    p, (audio, visual, caps), o_out, o_terms, o_grads = c2_oracle
    B, T, L, V = 128, 44, 24, 3201
    model = AVCaptioning(Vocab(V), 1.0, "none", device=dev, precision="bf16").to(dev)
    _load(model, p)
    out, _, _ = model(audio.to(dev), visual.to(dev), caps.to(dev))
    terms = Lm.ModalityWiseReconstructionLoss(out, caps.to(dev), **LAM)
    terms[0].mean().backward()
    torch.testing.assert_close(out.detach().cpu().double(), o_out, atol=5e-2, rtol=5e-2)
    assert float(terms[0]) == pytest.approx(o_terms[0], rel=2e-2)
    assert float(terms[1]) == pytest.approx(o_terms[1], rel=2e-2)
    assert float(terms[2]) == pytest.approx(o_terms[2], rel=2e-2)
    bad = []
    for k, v in model.named_parameters():
        c = cos(v.grad, o_grads[k])
        n1, n2 = float(v.grad.norm()), float(o_grads[k].norm())
        if c < 0.999 or abs(n1 - n2) > 5e-2 * n2 + 1e-9:
            bad.append(f"{k}: cosine {c:.5f}, norm {n1:.4e} vs {n2:.4e}")
    assert not bad, "\n".join(bad)
    # bf16 feature shards (the e2e input format) drive the same kernels: bit-identical log-probs
    with torch.no_grad():
        out_b, _, _ = model((audio.bfloat16()).to(dev), (visual.bfloat16()).to(dev), caps.to(dev))
    assert torch.equal(out_b, out.detach())


def test_c2_bf16_raw_features_vs_rounded_operand_oracle(dev):
    """The default bench.py line runs the C2 step on RAW synthetic features (audio 0..255, visual up to ~48), where
    rounding the operands to bf16 alone moves the gradients (see test_bf16_path_vs_fp64_oracle).  At the exact
    benchmarked shape the kernels must reproduce the fp64 oracle evaluated on the bf16-rounded weights and features:
    log-probs atol 5e-2, loss rtol 1e-3, LSTM / embedding / vocabulary gradients cosine >= 0.999, attention
    gradients (tanh.approx + saturated scores) cosine >= 0.9 (measured 0.945-0.98)."""
    from models import AVCaptioning
    import losses as Lm
    B, T, L, V = 128, 44, 24, 3201
    gen = torch.Generator().manual_seed(0)
    p = O.init_decoder_params("decoder.", 2176, V, gen=gen)
    audio, visual, caps = O.synth_batch(B, T, L, V, seed=1)
    pr = {k: v.bfloat16().double().requires_grad_() for k, v in p.items()}
    o, _, _ = O.av_forward(pr, audio.bfloat16().double(), visual.bfloat16().double(), caps, 1.0, "none", hoist=True)
    ot = O.modality_wise_loss(o, caps, **LAM)
    ot[0].backward()
    model = AVCaptioning(Vocab(V), 1.0, "none", device=dev, precision="bf16").to(dev)
    _load(model, p)
    out, _, _ = model(audio.to(dev), visual.to(dev), caps.to(dev))
    terms = Lm.ModalityWiseReconstructionLoss(out, caps.to(dev), **LAM)
    terms[0].mean().backward()
    torch.testing.assert_close(out.detach().cpu().double(), o.detach(), atol=5e-2, rtol=5e-2)
    assert float(terms[0]) == pytest.approx(float(ot[0]), rel=1e-3)
    bad = []
    for k, v in model.named_parameters():
        c = cos(v.grad, pr[k].grad)
        n1, n2 = float(v.grad.norm()), float(pr[k].grad.norm())
        att = ".attention." in k
        if c < (0.9 if att else 0.999) or abs(n1 - n2) > (0.2 if att else 5e-2) * n2 + 1e-9:
            bad.append(f"{k}: cosine {c:.5f}, norm {n1:.4e} vs {n2:.4e}")
    assert not bad, "\n".join(bad)


def test_c2_fp32_exact_path_vs_oracle(dev, c2_oracle):
    from models import AVCaptioning
    import losses as Lm
    p, (audio, visual, caps), o_out, o_terms, o_grads = c2_oracle
    V = 3201
    model = AVCaptioning(Vocab(V), 1.0, "none", device=dev, precision="fp32").to(dev)
    _load(model, p)
    out, _, _ = model(audio.to(dev), visual.to(dev), caps.to(dev))
    terms = Lm.ModalityWiseReconstructionLoss(out, caps.to(dev), **LAM)
    terms[0].mean().backward()
    torch.testing.assert_close(out.detach().cpu().double(), o_out, atol=1e-4, rtol=1e-4)
    assert float(terms[0]) == pytest.approx(o_terms[0], rel=1e-5)
    for k, v in model.named_parameters():
        ref = o_grads[k].float()
        torch.testing.assert_close(v.grad.cpu(), ref, rtol=1e-3, atol=1e-3 * float(ref.abs().max()) + 1e-9, msg=k)


def test_c3_fp32_greedy_ids_identical_to_oracle(dev, monkeypatch):
    """BASELINE configs[2] per-GPU shape: B=512, T=30, 30 positions, V=10547; exact caption agreement on the fp32 path."""
    from models import AVCaptioning
    B, T, L, V = 512, 30, 30, 10547
    gen = torch.Generator().manual_seed(0)
    p = O.init_decoder_params("decoder.", 2176, V, gen=gen)
    p["decoder.out.weight"] = p["decoder.out.weight"] * 8.0                    # peaky logits
    audio, visual, _ = O.synth_batch(B, T, L, V, seed=2, min_frames=10)
    model = AVCaptioning(Vocab(V), 0.0, "none", device=dev, precision="fp32").to(dev)
    _load(model, p)
    with torch.no_grad():
        ids = model.decoder.greedy_ids((audio.to(dev), visual.to(dev)), L).cpu()
        ref = O.av_greedy_ids(p, audio, visual, L)
    assert ids.shape == ref.shape == (B, L)
    same = (ids == ref).all(1)
    assert bool(same.all()), f"{int((~same).sum())} of {B} captions differ from the fp32 oracle"
    # the bf16 tensor-core path on the same inputs: agreement is REPORTED (rate), not required
    model.set_precision("bf16")
    with torch.no_grad():
        ids_b = model.decoder.greedy_ids((audio.to(dev), visual.to(dev)), L).cpu()
    rate = float((ids_b == ref).all(1).float().mean())
    print(f"bf16 greedy captions identical to fp32 oracle: {rate:.3f}")
    assert rate > 0.5
    # the opt-in projected-keys form of the decode step (MVC_B200_GREEDY_P=1, B > 148 rows: attention sums
    # P = keys . W_c^T rows, gate GEMM over h only) must tell the same story as the plain form
    monkeypatch.setenv("MVC_B200_GREEDY_P", "1")
    with torch.no_grad():
        ids_c = model.decoder.greedy_ids((audio.to(dev), visual.to(dev)), L).cpu()
    monkeypatch.delenv("MVC_B200_GREEDY_P")
    rate_c = float((ids_c == ref).all(1).float().mean())
    both = float((ids_c == ids_b).all(1).float().mean())
    print(f"projected-keys form: {rate_c:.3f}; == plain form: {both:.3f}")
    assert rate_c > 0.5 and rate_c >= rate - 0.05 and both > 0.5


def test_c2_backward_without_forward_side_preparation(dev, monkeypatch):
    """A training forward (save_for_backward = 1, recur2 path) prepares the backward pass's operands itself: transposed
    weights / keys under the persistent kernel, ctx rows and [emb ; ctx ; h]^T under the vocabulary projection.  A forward
    called with save_for_backward = 0 leaves all of that to mvc_decoder_backward: both orders must give the same
    gradients (same kernels on the same data; only the embedding scatter's atomic order may differ)."""
    from models import AVCaptioning
    from salstm import cabi
    import losses as Lm
    B, T, L, V = 128, 44, 24, 3201
    audio, visual, caps = O.synth_batch(B, T, L, V, seed=3)
    audio, visual, caps = (audio / 255.0).to(dev), (visual / 10.0).to(dev), caps.to(dev)

    def run():
        torch.manual_seed(11)
        model = AVCaptioning(Vocab(V), 1.0, "none", device=dev, precision="bf16").to(dev)
        out, _, _ = model(audio, visual, caps)
        Lm.ModalityWiseReconstructionLoss(out, caps, **LAM)[0].mean().backward()
        return out.detach().clone(), {k: v.grad.clone() for k, v in model.named_parameters()}

    out_a, g_a = run()
    lib = cabi.lib()
    orig = lib.mvc_decoder_forward
    seen = []

    def no_prep(*args):
        args = list(args)
        seen.append(args[13])
        args[13] = 0                                   # save_for_backward
        return orig(*args)

    monkeypatch.setattr(lib, "mvc_decoder_forward", no_prep)
    out_b, g_b = run()
    assert seen == [1]                                 # the autograd path asks for the preparation
    assert torch.equal(out_a, out_b)
    for k in g_a:
        torch.testing.assert_close(g_b[k], g_a[k], rtol=1e-5, atol=1e-7, msg=k)
