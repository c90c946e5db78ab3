"""CUDA-event timing of the segments of one C2 train step (forward / loss / backward / optimizer), warm, 20 steps."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "multimodal-video-captioning_b200"), ROOT]
import torch
import bench as Bn
from salstm import cabi
from salstm.trainer import FlatClipAdam
import losses as Lm
wl = sys.argv[1] if len(sys.argv) > 1 else "train"
dev = torch.device("cuda:0")
w = Bn.WORKLOADS[wl]
shape = Bn.SHAPES[w["shape"]]
model = Bn.build_model(wl, dev, "bf16")
bs = [tuple(t.to(dev) for t in b) for b in Bn.make_batches(shape, 4)]
loss_fn = Lm.ModalityWiseReconstructionLossBuilder(rec_type=w["rec"], **Bn.LAMBDAS)
opt = FlatClipAdam(model.parameters(), lr=1e-4)
ev = lambda: torch.cuda.Event(enable_timing=True)
acc = [0.0] * 4
N = 20
for i in range(N + 5):
    a, v, c = bs[i % 4]
    e = [ev() for _ in range(5)]
    opt.zero_grad()
    e[0].record()
    out, ar, vr = model(a, v, c)
    e[1].record()
    terms = loss_fn(out, c, a, ar, v, vr)
    e[2].record()
    terms[0].mean().backward()
    e[3].record()
    opt.step()
    e[4].record()
    torch.cuda.synchronize()
    if i >= 5:
        for k in range(4):
            acc[k] += e[k].elapsed_time(e[k + 1])
print(wl, "forward %.1f us | loss %.1f us | backward %.1f us | optimizer %.1f us | sum %.1f us (synchronised per step)" %
      tuple([x / N * 1e3 for x in acc] + [sum(acc) / N * 1e3]))
