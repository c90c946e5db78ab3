"""CPU-side checks of the drop-in surface: pickle identity of the module classes, a committed
whole-module pickle written by the UNMODIFIED reference (tools/make_golden_r2.py) loading into the repo's classes,
`from losses import ..., NLPScore` (train.py:12), the optimizer's torch.optim.Optimizer contract, the product-side
synthetic-batch generator, and the staged reference (oracle/_ref) used by `bench.py --impl reference`."""
import os
import pickle
import sys
import textwrap

import pytest
import torch

from conftest import GOLDEN, PKG, ROOT
from oracle import salstm_oracle as O


def test_class_module_paths_match_reference():
    import models
    from models.captioning import AVCaptioning, AVCaptioningDual
    from models.features_captioning import FeaturesCaptioning
    from models.reconstructor import GlobalReconstructor, LocalReconstructor
    from models.temporal_attention import TemporalAttention
    want = {AVCaptioning: "models.captioning", AVCaptioningDual: "models.captioning",
            FeaturesCaptioning: "models.features_captioning", GlobalReconstructor: "models.reconstructor",
            LocalReconstructor: "models.reconstructor", TemporalAttention: "models.temporal_attention"}
    for cls, mod in want.items():
        assert cls.__module__ == mod
        assert pickle.loads(pickle.dumps(cls)) is cls          # pickle resolves the path back to the same class
    assert models.AVCaptioning is AVCaptioning and os.path.dirname(models.__file__) == os.path.join(PKG, "models")


def test_reference_pickle_loads_into_repo_classes():
    """torch.save(model) written by the reference (train.py:162-173) -> our classes, with a working default precision."""
    import models  # noqa: F401  (PKG is first on sys.path: the pickle's `models.*` paths resolve to the repo's package)
    from salstm import modules as M
    blob = torch.load(os.path.join(GOLDEN, "ref_decoder_tiny.pt"), weights_only=False)
    dec, grec, lrec = blob["decoder"], blob["global"], blob["local"]
    assert type(dec) is M.FeaturesCaptioning and type(grec) is M.GlobalReconstructor and type(lrec) is M.LocalReconstructor
    assert type(dec.attention) is M.TemporalAttention
    assert "precision" not in dec.__dict__ and dec.precision in ("fp32", "bf16")     # class-level default
    assert dec._dims(5, 6, 7)[:8] == (5, 6, 16, 24, 12, 8, 23, 7)
    assert grec._dims(5, 7, 0)[:4] == (5, 7, 24, 16) and lrec._dims(5, 7, 6)[4] == 8
    assert sorted(dec.state_dict()) == sorted(["embedding.weight", "attention.b", "attention.W.weight",
                                               "attention.U.weight", "attention.w.weight", "rnn.weight_ih_l0",
                                               "rnn.weight_hh_l0", "rnn.bias_ih_l0", "rnn.bias_hh_l0", "out.weight",
                                               "out.bias"])


def test_repo_pickle_roundtrip(tmp_path):
    from models import AVCaptioningDual

    class V:
        stoi = {"<SOS>": 1, "<EOS>": 2}

        def __len__(self):
            return 11

    m = AVCaptioningDual.__new__(AVCaptioningDual)       # avoid allocating the 34 M-parameter model: path check only
    torch.nn.Module.__init__(m)
    m.vocab_size = 11
    data = pickle.dumps(m)
    assert b"models.captioning" in data and b"salstm.modules" not in data
    assert type(pickle.loads(data)) is AVCaptioningDual


def test_nlpscore_matches_the_reference_scorers():
    """`from losses import ..., NLPScore` (train.py:12) resolves to this package, and BLEU-1..4 / ROUGE-L / CIDEr equal
    what the reference's pycocoevalcap scorers return on the committed seeded sentence sets (tools/make_golden_r2.py)."""
    import json
    import math
    import losses as L
    assert os.path.dirname(L.__file__) == PKG
    from losses import ModalityWiseReconstructionLossBuilder, NLPScore  # noqa: F401
    cases = json.load(open(os.path.join(GOLDEN, "nlp_scores_small.json")))
    assert len(cases) >= 5
    for c in cases:
        s = NLPScore(c["gts"], c["res"])
        assert sorted(s) == ["Bleu_1", "Bleu_2", "Bleu_3", "Bleu_4", "CIDEr", "METEOR", "ROUGE_L"]
        for k in range(4):
            assert s[f"Bleu_{k + 1}"] == pytest.approx(c["Bleu"][k], abs=1e-12)
        assert s["ROUGE_L"] == pytest.approx(c["ROUGE_L"], abs=1e-12)
        assert s["CIDEr"] == pytest.approx(c["CIDEr"], abs=1e-12)
        assert math.isnan(s["METEOR"]) or 0.0 <= s["METEOR"] <= 1.0      # Java scorer: delegated when available
    with pytest.raises(AssertionError):
        NLPScore({"a": ["x"]}, {"b": ["x"]})                              # same contract as the reference scorers


def test_flat_clip_adam_is_a_torch_optimizer():
    from salstm.trainer import FlatClipAdam
    net = torch.nn.Linear(3, 2)
    opt = FlatClipAdam(net.parameters(), lr=1e-3)
    assert isinstance(opt, torch.optim.Optimizer) and len(opt.param_groups) == 1
    sched = torch.optim.lr_scheduler.ReduceLROnPlateau(opt, mode="min", factor=0.5, patience=0, min_lr=1e-7)   # train.py:90-97
    sched.step(1.0); sched.step(2.0)
    assert opt.lr == pytest.approx(5e-4) and opt.param_groups[0]["lr"] == pytest.approx(5e-4)
    sd = opt.state_dict()
    assert sd["step"] == 0 and sd["param_group"]["lr"] == pytest.approx(5e-4)
    with pytest.raises(NotImplementedError):
        FlatClipAdam(net.parameters(), amsgrad=False)


def test_grad_arena_refuses_aliasing():
    """the arena view may be handed to a backward kernel only when autograd will ADOPT it."""
    from salstm import functional as Fn
    p = torch.nn.Parameter(torch.zeros(4))
    view = torch.zeros(4)
    arena = Fn.GradArena([p], [view])
    t = p.detach()
    got = arena.take(t)
    assert got is not None and got.data_ptr() == view.data_ptr()
    assert arena.take(t) is None                      # second node of the same backward: fresh tensor instead
    arena.new_step()
    p.grad = torch.ones(4)                            # accumulation / zero_grad(set_to_none=False)
    assert arena.take(t) is None
    p.grad = None
    assert arena.take(t) is not None


def test_product_synth_generator_matches_the_oracles():
    from salstm.synth import frame_lengths, synth_batch
    for kw in (dict(B=5, T=7, L=6, V=31, seed=3), dict(B=9, T=4, L=9, V=50, Fa=8, Fv=16, seed=11, min_frames=2, min_cap=3)):
        a, b = synth_batch(**kw), O.synth_batch(**kw)
        assert all(torch.equal(x, y) for x, y in zip(a, b))
        n = frame_lengths(a[0], a[1])
        T = a[0].shape[1]
        pad = torch.arange(T).unsqueeze(0) >= n.unsqueeze(1)
        assert float(a[0][pad].abs().sum()) == 0 and float(a[1][pad].abs().sum()) == 0 and int(n.min()) >= 1


def test_staged_reference_loads_and_agrees_with_the_oracle():
    """oracle/_ref (byte copies of the reference's hot-path files, staged by oracle/build_ref.py; what
    `bench.py --impl reference` times on the GPU box) gives the oracle's numbers on a small case."""
    from oracle import build_ref
    if not build_ref.available():
        pytest.skip("oracle/_ref not staged (needs /root/reference at build time)")
    ref = build_ref.load()
    import models
    assert os.path.dirname(models.__file__) == os.path.join(PKG, "models")       # the product's package is untouched
    assert ref.AVCaptioning.__module__ == "models.captioning" and ref.AVCaptioning is not models.AVCaptioning
    torch.manual_seed(0)
    dec = ref.FeaturesCaptioning(in_feature_size=10, output_size=13, rnn_hidden_size=8, embedding_size=6, attn_size=4)
    feats = torch.rand(3, 5, 10)
    _, _, caps = O.synth_batch(3, 5, 6, 13, Fa=1, Fv=1, seed=2, min_cap=3)
    with torch.no_grad():
        out, hid = dec.decode(feats, caps, 6, 1.0)
        o_out, o_hid = O.decoder_decode({k: v for k, v in dec.state_dict().items()}, "", feats, caps, 6, 1.0)
    torch.testing.assert_close(out, o_out, atol=1e-6, rtol=1e-5)
    torch.testing.assert_close(hid, o_hid, atol=1e-6, rtol=1e-5)


def test_launcher_puts_the_package_ahead_of_the_script_directory(tmp_path):
    """`python src/train.py` resolves `models` to the script's own directory whatever PYTHONPATH says;
    salstm/launch.py runs the unmodified script with the package first and the script's directory behind it."""
    import subprocess
    src = tmp_path / "ref" / "src"
    (src / "models").mkdir(parents=True)
    (src / "models" / "__init__.py").write_text("WHO = 'reference'\n")
    (src / "losses.py").write_text("def NLPScore(ref, hypo):\n    return {'CIDEr': 1.5}\n")
    (src / "get_loader.py").write_text("WHO = 'reference'\n")
    (src / "train.py").write_text(textwrap.dedent("""
        import sys, os
        from get_loader import WHO
        from losses import ModalityWiseReconstructionLossBuilder, NLPScore
        from models import AVCaptioning, AVCaptioningDual
        import models, losses
        sc = NLPScore({"v": ["a b c"]}, {"v": ["a b c"]})
        ref = losses._reference_losses()
        print("RESULT", WHO, os.path.dirname(models.__file__), os.path.dirname(losses.__file__), round(sc["Bleu_1"], 6),
              ref.NLPScore({}, {}), sys.argv[1:])
    """))
    out = subprocess.run([sys.executable, os.path.join(PKG, "salstm", "launch.py"), str(src / "train.py"), "--gpu", "0"],
                         capture_output=True, text=True, cwd=str(tmp_path), timeout=300)
    assert out.returncode == 0, out.stderr
    line = [ln for ln in out.stdout.splitlines() if ln.startswith("RESULT")][0]
    # models / losses from this package, get_loader from the script directory, and the reference's own losses.py (METEOR
    # delegate) discoverable behind the package
    assert line == f"RESULT reference {os.path.join(PKG, 'models')} {PKG} 1.0 {{'CIDEr': 1.5}} ['--gpu', '0']"


def test_vectorised_detokenisation_equals_decode_indexes():
    """decode_batch(vocab, ids) == [vocab.decode_indexes(row[1:]) ...] (get_loader.py:79-89), including rows
    without <EOS>, rows that start with <EOS>, and vocabularies without a dense table (fallback)."""
    from salstm.modules import decode_batch

    class V:
        def __init__(self, n):
            self.itos = {0: "<PAD>", 1: "<SOS>", 2: "<EOS>", 3: "<UNK>"}
            self.itos.update({i: f"w{i}" for i in range(4, n)})

        def decode_indexes(self, idx):
            return O.decode_indexes(self.itos, idx)

    g = torch.Generator().manual_seed(0)
    v = V(40)
    ids = torch.randint(0, 40, (64, 12), generator=g)
    ids[0, 1:] = torch.randint(3, 40, (11,), generator=g)       # no EOS at all
    ids[1, 1] = 2                                               # EOS first -> empty caption
    want = [v.decode_indexes(r[1:]) for r in ids.tolist()]
    assert decode_batch(v, ids.tolist()) == want and decode_batch(v, ids.numpy()) == want
    assert want[1] == "" and "<EOS>" not in " ".join(want)

    class Sparse(V):
        def __init__(self):
            super().__init__(40)
            del self.itos[17]
    s = Sparse()
    ok = ids.clone(); ok[ok == 17] = 5
    assert decode_batch(s, ok.tolist()) == [s.decode_indexes(r[1:]) for r in ok.tolist()]


def test_flat_clip_adam_state_dict_round_trip_is_layout_independent():
    """The flat buffers pad every parameter to a multiple of 4 elements (16-byte aligned views for the float4 kernels);
    state_dict stores the moments per live parameter without the pads and load_state_dict scatters them back."""
    from salstm.trainer import FlatClipAdam
    torch.manual_seed(0)

    def make():
        return [torch.nn.Parameter(torch.randn(3, 5)), torch.nn.Parameter(torch.randn(7)),
                torch.nn.Parameter(torch.randn(2, 2)), torch.nn.Parameter(torch.randn(9))]

    ps = make()
    for p in ps[:3]:                                   # the last parameter never gets a gradient: not live
        p.grad = torch.randn_like(p)
    opt = FlatClipAdam(ps, lr=1e-3)
    ranges = opt._sync_views()
    assert opt._offsets == [0, 16, 24] and opt._n_flat == 28 and ranges == [(0, 28)]
    assert all(p.data.data_ptr() % 16 == 0 for p in ps[:3])
    assert float(opt.flat_p[15]) == 0.0 and float(opt.flat_g[23]) == 0.0          # pads are zero
    for buf in (opt.m, opt.v, opt.vmax):
        buf.copy_(torch.randn(buf.numel()))
    opt.step_count = 7
    sd = opt.state_dict()
    assert sd["exp_avg"].numel() == 15 + 7 + 4 and sd["live"] == [0, 1, 2]
    qs = make()
    opt2 = FlatClipAdam(qs, lr=5e-4)
    opt2.load_state_dict(sd)
    assert opt2.step_count == 7 and opt2.lr == pytest.approx(1e-3)
    for k, (p, o) in enumerate(zip(opt._live, opt._offsets)):
        n = p.numel()
        for a, b in ((opt.m, opt2.m), (opt.v, opt2.v), (opt.vmax, opt2.vmax)):
            assert torch.equal(a[o:o + n], b[opt2._offsets[k]:opt2._offsets[k] + n])
    assert qs[3].grad is None
