"""Phase timing of the persistent backward recurrence kernel (CTA 0, SM clock) over one train step."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "multimodal-video-captioning_b200"), ROOT]
import torch
import bench as Bn
from salstm import cabi
import losses as Lm
dev = torch.device("cuda:0")
lib = cabi.lib()
shape = Bn.SHAPES["msvd"]
B, T, L, V = shape
model = Bn.build_model("train", dev, "bf16")
b = [t.to(dev) for t in Bn.make_batches(shape, 1)[0]]
loss_fn = Lm.ModalityWiseReconstructionLossBuilder(rec_type="none", **Bn.LAMBDAS)
S = L - 1
def step():
    out, ar, vr = model(b[0], b[1], b[2])
    loss_fn(out, b[2], b[0], ar, b[1], vr)[0].backward()
for _ in range(3):
    step()
buf = torch.zeros(10 * S, dtype=torch.int64, device=dev)
lib.mvc_debug_set_recur_bwd_prof(cabi.ptr(buf))
step()
torch.cuda.synchronize()
lib.mvc_debug_set_recur_bwd_prof(None)
t = buf.cpu().view(S, 10).double()
names = ["P2 GEMM + park", "cluster barrier", "reduce -> dxh", "grid barrier 1", "P3 attention bwd", "cluster barrier (dwq)", "P4 dh + cell bwd", "grid barrier 2"]
d = (t[:, 1:9] - t[:, 0:8]) / 1.9e3
print("per-step phase durations (us, mean over steps 2..S):")
for i, n in enumerate(names):
    print(f"  {n:28s} {d[2:, i].mean():7.2f}  (min {d[2:, i].min():6.2f} max {d[2:, i].max():6.2f})")
print(f"  {'step total':28s} {((t[2:, 8] - t[2:, 0]) / 1.9e3).mean():7.2f}")
