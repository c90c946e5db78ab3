import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "multimodal-video-captioning_b200"), ROOT, os.path.join(ROOT, "tests")]
import torch
from oracle import salstm_oracle as O
from test_oracle_golden import _wrapper_params
from test_gpu_parity import Vocab, _load, cos
from models import AVCaptioning
import losses as L
dev = torch.device("cuda:0")
B, T, Lc, V = 6, 9, 8, 211
p = _wrapper_params("joint", V, "none", 77)
audio, visual, caps = O.synth_batch(B, T, Lc, V, seed=5, min_frames=2, min_cap=4)
res = {}
for prec in ("fp32", "bf16"):
    model = AVCaptioning(Vocab(V), 1.0, "none", device=dev, precision=prec).to(dev)
    _load(model, p)
    out, _, _ = model(audio.to(dev), visual.to(dev), caps.to(dev))
    terms = L.ModalityWiseReconstructionLoss(out, caps.to(dev), None, None, None, None, 0.0005, 0, 0, "none")
    terms[0].backward()
    res[prec] = ({k: v.grad.clone() for k, v in model.named_parameters()}, out.detach())
g32, g16 = res["fp32"][0], res["bf16"][0]
print("out diff", (res["fp32"][1] - res["bf16"][1]).abs().max().item())
for k in g32:
    print(k, "cos", round(cos(g32[k], g16[k]), 5), "norm", g32[k].norm().item(), g16[k].norm().item())
E = 300
k = "decoder.rnn.weight_ih_l0"
print("w_ih emb part cos", cos(g32[k][:, :E], g16[k][:, :E]), g32[k][:, :E].norm().item(), g16[k][:, :E].norm().item())
print("w_ih ctx part cos", cos(g32[k][:, E:], g16[k][:, E:]), g32[k][:, E:].norm().item(), g16[k][:, E:].norm().item())
print("w_ih audio part cos", cos(g32[k][:, E:E+128], g16[k][:, E:E+128]))
print("w_ih visual part cos", cos(g32[k][:, E+128:], g16[k][:, E+128:]))
for gate, nm in enumerate("ifgo"):
    sl = slice(gate * 512, (gate + 1) * 512)
    print("gate", nm, cos(g32[k][sl, E:], g16[k][sl, E:]), g32[k][sl, E:].norm().item(), g16[k][sl, E:].norm().item())
