"""Pin the CPU oracle (oracle/salstm_oracle.py) against outputs of the
unmodified reference modules (tests/golden/*.npz, made by tools/make_golden.py).
fp32 tolerances: log-probs/hiddens atol 1e-4 (the oracle uses a closed-form
LSTM cell and its own op order; the reference calls oneDNN), ids identical."""
import numpy as np
import pytest
import torch

from conftest import load_golden, sub
from oracle import salstm_oracle as O

TOL = dict(atol=1e-4, rtol=1e-4)


def close(a, b, **kw):
    kw = {**TOL, **kw}
    torch.testing.assert_close(a, b, **kw)


def test_attention_matches_reference():
    g = load_golden("attention_small")
    p = sub(g, "p.")
    ctx, al = O.soft_attention(p, "", g["q"], g["k"])
    close(ctx, g["ctx"]); close(al, g["alpha"])
    ctx, al = O.soft_attention(p, "", g["q"], g["k"], mask=g["mask"])
    close(ctx, g["ctx_masked"]); close(al, g["alpha_masked"])
    assert (al[~g["mask"]] == 0).all()
    # hoisting U.k out of the loop is exact algebra
    uk = g["k"] @ p["U.weight"].t()
    ctx2, _ = O.soft_attention(p, "", g["q"], g["k"], mask=g["mask"], keys_proj=uk)
    close(ctx2, g["ctx_masked"])


@pytest.mark.parametrize("aten", [False, True])
def test_decoder_step_and_teacher_forced(aten):
    g = load_golden("decoder_small")
    p = sub(g, "p.")
    lp, h1, c1, al = O.decoder_step(p, "", g["feats"], g["step_h0"], g["step_c0"], g["step_w0"], aten_lstm=aten)
    close(lp, g["step_lp"]); close(h1, g["step_h1"]); close(c1, g["step_c1"]); close(al, g["step_alpha"])
    L = g["caps"].shape[0]
    out, hid = O.decoder_decode(p, "", g["feats"], g["caps"], L, 1.0, aten_lstm=aten, hoist=not aten)
    close(out, g["tf1_out"]); close(hid, g["tf1_hid"])
    assert out[0].abs().max() == 0 and hid[0].abs().max() == 0


def test_decoder_gradients():
    g = load_golden("decoder_small")
    p = {k: v.clone().requires_grad_() for k, v in sub(g, "p.").items()}
    L, V = g["caps"].shape[0], g["tf1_out"].shape[2]
    out, hid = O.decoder_decode(p, "", g["feats"], g["caps"], L, 1.0, hoist=True)
    loss = torch.nn.functional.nll_loss(out[1:].reshape(-1, V), g["caps"][1:].reshape(-1), ignore_index=0) \
        + 0.01 * hid.pow(2).sum()
    close(loss, g["tf1_loss"], rtol=1e-5)
    loss.backward()
    for k, v in p.items():
        close(v.grad, g["tf1_grad." + k], atol=2e-5, rtol=1e-3)


def test_decoder_sampling_and_greedy():
    g = load_golden("decoder_small")
    p = sub(g, "p.")
    L = g["caps"].shape[0]
    torch.manual_seed(1234)
    flags = O.teacher_flags(g["caps"], L, 0.5)
    assert flags == [bool(x) for x in g["tf05_flags"].tolist()]
    assert len(flags) == L - 1          # reference draws exactly L-1 numbers
    out, hid = O.decoder_decode(p, "", g["feats"], g["caps"], L, 0.5, flags=flags)
    close(out, g["tf05_out"]); close(hid, g["tf05_hid"])
    out, hid = O.decoder_decode(p, "", g["feats"], g["caps"], L, 0.0)
    close(out, g["tf0_out"]); close(hid, g["tf0_hid"])
    out, hid = O.decoder_decode(p, "", g["feats"], None, 12, hoist=True)
    close(out, g["greedy_out"]); close(hid, g["greedy_hid"])
    assert torch.equal(out.argmax(2).t(), g["greedy_ids"])
    s0 = torch.random.get_rng_state()
    O.teacher_flags(None, 12, 0.5)      # captions=None draws nothing
    assert torch.equal(s0, torch.random.get_rng_state())


def test_batched_rng_draw_equals_reference_draws():
    """The module layer draws its teacher-forcing flags with one torch.rand(n); the reference draws n times
    torch.rand(1) (features_captioning.py:116).  Values and final generator state must be identical."""
    import sys
    from conftest import PKG
    if PKG not in sys.path:
        sys.path.insert(0, PKG)
    from salstm.functional import teacher_flags
    for n, ratio in ((2, 0.5), (9, 0.3), (24, 0.7), (31, 1.0), (65, 0.5)):
        caps = torch.ones(n, 1, dtype=torch.long)
        torch.manual_seed(77)
        ref = O.teacher_flags(caps, n, ratio)
        s_ref = torch.random.get_rng_state()
        torch.manual_seed(77)
        got = teacher_flags(caps, n, ratio)
        assert got == ref and torch.equal(s_ref, torch.random.get_rng_state())


def _prefix(ids):
    ids = [int(x) for x in ids]
    return ids[: ids.index(O.EOS) + 1] if O.EOS in ids[1:] else ids


def test_beam_search_prefix_to_eos():
    g = load_golden("decoder_small")
    p = sub(g, "p.")
    ids = O.decoder_beam_search(p, "", g["feats"], max_len=10, width=3)
    assert ids.shape == g["beam_ids"].shape == (g["feats"].shape[0], 12)
    for a, b in zip(ids, g["beam_ids"]):
        assert _prefix(a) == _prefix(b)


def test_reconstructors_and_losses():
    g = load_golden("recon_loss_small")
    hid = g["hid"].clone().requires_grad_()
    pg = {k: v.clone().requires_grad_() for k, v in sub(g, "g.").items()}
    rec = O.global_reconstruct(pg, "", hid, g["outs"], g["caps"])
    close(rec, g["g_rec"])
    loss = O.global_recon_loss(g["feats"], rec, g["caps"] != 0)
    close(loss, g["g_loss"], rtol=1e-5)
    loss.backward()
    close(hid.grad, g["g_dhid"], atol=1e-6, rtol=1e-3)
    for k, v in pg.items():
        close(v.grad, g["g_grad." + k], atol=1e-6, rtol=1e-3)
    close(O.global_reconstruct(pg, "", hid, g["outs"], None), g["g_rec_nocap"])

    hid = g["hid"].clone().requires_grad_()
    pl = {k: v.clone().requires_grad_() for k, v in sub(g, "l.").items()}
    T = g["feats"].shape[1]
    for hoist in (False, True):
        rec = O.local_reconstruct(pl, "", hid, g["outs"], g["caps"], T, hoist=hoist)
        close(rec, g["l_rec"])
    loss = O.local_recon_loss(g["feats"], rec)
    close(loss, g["l_loss"], rtol=1e-5)
    loss.backward()
    close(hid.grad, g["l_dhid"], atol=1e-6, rtol=1e-3)
    for k, v in pl.items():
        close(v.grad, g["l_grad." + k], atol=1e-6, rtol=1e-3)

    lo = g["outs"].clone().requires_grad_()
    terms = O.modality_wise_loss(lo, g["caps"], g["loss_afeat"], g["loss_arec"], g["feats"], g["loss_vrec"],
                                 reg_lambda=0.0005, audio_recon_lambda=0.00005, visual_recon_lambda=0.5,
                                 rec_type="global")
    close(torch.stack([t.detach() for t in terms]), g["loss_terms"], rtol=1e-5)
    terms[0].backward()
    close(lo.grad, g["loss_dout"], atol=1e-7, rtol=1e-4)


def _wrapper_params(kind, V, rec_type, seed):
    gen = torch.Generator().manual_seed(seed)
    p = {}
    if kind == "joint":
        p.update(O.init_decoder_params("decoder.", 2176, V, gen=gen))
        if rec_type != "none":
            p.update(O.init_recon_params("reconstructor.", rec_type, 512, 2176, gen=gen))
        p["decoder.out.weight"] *= 6.0
    else:
        p.update(O.init_decoder_params("v_decoder.", 2048, V, gen=gen))
        p.update(O.init_decoder_params("a_decoder.", 128, V, gen=gen))
        if rec_type != "none":
            p.update(O.init_recon_params("v_reconstructor.", rec_type, 512, 2048, gen=gen))
            p.update(O.init_recon_params("a_reconstructor.", rec_type, 512, 128, gen=gen))
        p["v_decoder.out.weight"] *= 6.0
        p["a_decoder.out.weight"] *= 6.0
    return p


@pytest.mark.parametrize("kind", ["joint", "dual"])
@pytest.mark.parametrize("rec_type", ["none", "global", "local"])
def test_full_width_wrappers(kind, rec_type):
    g = load_golden(f"wrapper_{kind}_{rec_type}")
    B, T, L, V = (int(g[k]) for k in "BTLV")
    p = {k: v.requires_grad_() for k, v in _wrapper_params(kind, V, rec_type, int(g["seed"])).items()}
    audio, visual, caps = O.synth_batch(B, T, L, V, seed=int(g["data_seed"]), min_frames=2, min_cap=4)
    fwd = O.av_forward if kind == "joint" else O.av_dual_forward
    out, arec, vrec = fwd(p, audio, visual, caps, 1.0, rec_type, hoist=True)
    close(out, g["out"], atol=2e-4)
    if rec_type != "none":
        close(arec, g["arec"], atol=2e-4); close(vrec[:, :, ::16], g["vrec"], atol=2e-4)
    terms = O.modality_wise_loss(out, caps, audio, arec, visual, vrec, reg_lambda=0.0005,
                                 audio_recon_lambda=0.00005, visual_recon_lambda=0.5, rec_type=rec_type)
    close(torch.stack([t.detach() for t in terms]), g["loss_terms"], rtol=2e-5, atol=1e-6)
    terms[0].backward()
    for k, v in p.items():
        if "gnorm." + k not in g:
            continue
        close(v.grad.norm(), g["gnorm." + k], rtol=2e-3, atol=1e-7)
        flat = v.grad.flatten()
        ref = g["gslice." + k]
        close(flat[:: max(1, flat.numel() // 64)][:64], ref, rtol=5e-3, atol=2e-6 + 1e-3 * float(ref.abs().max()))
    with torch.no_grad():
        out0, _, _ = fwd(p, audio, visual, caps, 0.0, rec_type, hoist=True)
        close(out0, g["out_tf0"], atol=2e-4)
        if rec_type == "none":
            itos = {0: "<PAD>", 1: "<SOS>", 2: "<EOS>", 3: "<UNK>", **{i: f"w{i}" for i in range(4, V)}}
            ids = (O.av_greedy_ids if kind == "joint" else O.av_dual_greedy_ids)(p, audio, visual, 10)
            assert [O.decode_indexes(itos, r[1:]) for r in ids] == g["greedy_txt"].tolist()
            if kind == "joint":
                assert torch.equal(ids, g["greedy_ids"])
                feats = torch.cat([audio, visual], -1)
                b = O.decoder_beam_search(p, "decoder.", feats, max_len=10, width=3)
                assert [O.decode_indexes(itos, r[1:]) for r in b] == g["beam_txt"].tolist()
