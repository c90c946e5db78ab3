"""CPU-side checks of the drop-in boundary: the C-ABI library loads without a GPU and
exports every symbol include/mvc_b200.h declares; the Python binding table covers the same
list; the product path refuses to run without CUDA (no CPU fallback)."""
import os
import re

import pytest
import torch

from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "mvc_b200.h")


def header_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mvc_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as g
    g.build()
    from salstm import cabi
    return cabi


def test_library_exports_every_header_symbol(built):
    lib = built.load()
    names = header_symbols()
    assert len(names) >= 40
    for n in names:
        assert hasattr(lib, n), f"libmvc_b200.so does not export {n}"
    assert lib.mvc_version() >= 100


def test_binding_table_matches_header(built):
    assert sorted(built.SIGNATURES) == header_symbols()


def test_workspace_queries_need_no_gpu(built):
    import ctypes as C
    lib = built.load()
    d = built.DecoderDims(128, 44, 2176, 512, 300, 256, 3201, 24, built.BF16)
    fwd = lib.mvc_decoder_fwd_workspace_bytes(C.byref(d), 1)
    bwd = lib.mvc_decoder_bwd_workspace_bytes(C.byref(d))
    assert 50e6 < fwd < 2e9 and 50e6 < bwd < 4e9
    r = built.ReconDims(128, 24, 512, 2176, 256, 44, built.BF16)
    assert lib.mvc_global_recon_workspace_bytes(C.byref(r)) > 0
    assert lib.mvc_local_recon_workspace_bytes(C.byref(r)) > 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_product_path_fails_loudly_without_cuda(built):
    from salstm.modules import FeaturesCaptioning
    dec = FeaturesCaptioning(in_feature_size=8, output_size=11, rnn_hidden_size=8, embedding_size=8, attn_size=8)
    with pytest.raises(RuntimeError, match="no CUDA device|CUDA only"):
        dec.decode(torch.zeros(2, 3, 8), torch.ones(4, 2, dtype=torch.int64), 4, 1.0)


def test_unsupported_configs_raise():
    from salstm.modules import FeaturesCaptioning, GlobalReconstructor
    with pytest.raises(NotImplementedError):
        FeaturesCaptioning(in_feature_size=8, output_size=11, rnn_type="GRU")
    with pytest.raises(NotImplementedError):
        GlobalReconstructor(decoder_size=8, hidden_size=8, rnn_num_layers=2)


def test_state_dict_names_match_reference():
    """Key names / shapes listed in SURVEY.md §8a (a2, a8, a9): old checkpoints must load."""
    from models import AVCaptioning, AVCaptioningDual

    class V:
        stoi = {"<SOS>": 1, "<EOS>": 2}

        def __len__(self):
            return 3201

    m = AVCaptioning(V(), reconstructor_type="local")
    sd = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert sd["decoder.embedding.weight"] == (3201, 300)
    assert sd["decoder.attention.U.weight"] == (256, 2176)
    assert sd["decoder.attention.b"] == (256,)
    assert sd["decoder.rnn.weight_ih_l0"] == (2048, 2476)
    assert sd["decoder.out.weight"] == (3201, 512)
    assert sd["reconstructor.rnn.weight_hh_l0"] == (8704, 2176)
    assert sd["reconstructor.attention.W.weight"] == (256, 2176)
    assert sum(v.numel() for v in m.decoder.parameters()) == 9414573
    d = AVCaptioningDual(V(), reconstructor_type="global")
    names = set(d.state_dict())
    assert {"v_decoder.out.bias", "a_decoder.rnn.bias_hh_l0", "output_fc.weight", "v_reconstructor.rnn.weight_ih_l0",
            "a_reconstructor.rnn.weight_hh_l0"} <= names
    assert d.a_reconstructor.rnn.weight_ih_l0.shape == (512, 1024)
