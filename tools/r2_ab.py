"""A/B of the two persistent-kernel generations against the fp64 oracle (run on the GPU box):
    MVC_B200_RECUR2=0 python tools/r2_ab.py raw     # first generation
    python tools/r2_ab.py raw                       # projected keys (recur2)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "multimodal-video-captioning_b200"), ROOT, os.path.join(ROOT, "tests")]
import torch
import __graft_entry__ as g
g.build()
from oracle import salstm_oracle as O
from models import AVCaptioning
import losses as L

mode = sys.argv[1] if len(sys.argv) > 1 else "raw"
B, T, Lc, V = (int(x) for x in (sys.argv[2:6] if len(sys.argv) >= 6 else (128, 44, 24, 3201)))
dev = torch.device("cuda:0")


class Vocab:
    stoi = {"<SOS>": 1, "<EOS>": 2}
    def __len__(self): return V


gen = torch.Generator().manual_seed(0)
p = O.init_decoder_params("decoder.", 2176, V, gen=gen)
audio, visual, caps = O.synth_batch(B, T, Lc, V, seed=1)
if mode == "unit":
    audio, visual = audio / 255.0, visual / 10.0
model = AVCaptioning(Vocab(), 1.0, "none", device=dev, precision="bf16").to(dev)
sd = model.state_dict(); sd.update({k: p[k].clone() for k in sd if k in p}); model.load_state_dict(sd)
out, _, _ = model(audio.to(dev), visual.to(dev), caps.to(dev))
t = L.ModalityWiseReconstructionLoss(out, caps.to(dev), reg_lambda=0.0005)
t[0].backward()
pd = {k: v.double().requires_grad_() for k, v in p.items()}
o, _, _ = O.av_forward(pd, audio.double(), visual.double(), caps, 1.0, "none", hoist=True)
ot = O.modality_wise_loss(o, caps, reg_lambda=0.0005)
ot[0].backward()
print("recur2" if os.environ.get("MVC_B200_RECUR2", "1") != "0" else "recur1", mode, "B,T,L,V", B, T, Lc, V)
print("max |dlogp|", float((out.detach().cpu().double() - o.detach()).abs().max()), "loss", float(t[0]), float(ot[0]))
for k, v in model.named_parameters():
    a, b = v.grad.detach().cpu().double().flatten(), pd[k].grad.flatten()
    print(f"{k:34s} cos {float(a @ b / (a.norm() * b.norm() + 1e-30)):.5f}  norm {float(a.norm()):.4e} / {float(b.norm()):.4e}")
# how much of the deviation is the bf16 rounding of the operands alone?  fp64 oracle on bf16-rounded weights + features
pr = {k: v.bfloat16().double().requires_grad_() for k, v in p.items()}
o2, _, _ = O.av_forward(pr, audio.bfloat16().double(), visual.bfloat16().double(), caps, 1.0, "none", hoist=True)
O.modality_wise_loss(o2, caps, reg_lambda=0.0005)[0].backward()
print("-- fp64 oracle with bf16-rounded operands vs exact:  max |dlogp|", float((o2.detach() - o.detach()).abs().max()))
for k in pr:
    a, b = pr[k].grad.flatten(), pd[k].grad.flatten()
    c = dict(model.named_parameters())[k.replace("decoder.", "decoder.")].grad.detach().cpu().double().flatten()
    print(f"{k:34s} cos(rounded,exact) {float(a @ b / (a.norm() * b.norm() + 1e-30)):.5f}   cos(kernel,rounded) {float(a @ c / (a.norm() * c.norm() + 1e-30)):.5f}")
