"""Phase timing of the persistent recurrence kernel (CTA 0, SM clock) over one teacher-forced forward."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "multimodal-video-captioning_b200"), ROOT]
import torch
import bench as Bn
from salstm import cabi
dev = torch.device("cuda:0")
lib = cabi.lib()
shape = Bn.SHAPES["msvd"]
B, T, L, V = shape
model = Bn.build_model("train", dev, "bf16")
b = [t.to(dev) for t in Bn.make_batches(shape, 1)[0]]
S = L - 1
for _ in range(3):
    model(b[0], b[1], b[2])
buf = torch.zeros(10 * S, dtype=torch.int64, device=dev)
lib.mvc_debug_set_recur_prof(cabi.ptr(buf))
with torch.no_grad():
    model(b[0], b[1], b[2])
torch.cuda.synchronize()
lib.mvc_debug_set_recur_prof(None)
t = buf.cpu().view(S, 10).double()
names = ["A1-A3 wq + cluster barrier", "scores + softmax", "A6 ctx (key ring)", "grid barrier 1", "MMA + park", "cluster barrier 2", "reduce + cell", "grid barrier 2"]
d = (t[:, 1:9] - t[:, 0:8]) / 1.9e3   # us at ~1.9 GHz
print("per-step phase durations (us, mean over steps 2..S):")
for i, n in enumerate(names):
    print(f"  {n:30s} {d[2:, i].mean():7.2f}  (min {d[2:, i].min():6.2f} max {d[2:, i].max():6.2f})")
print(f"  {'step total':30s} {((t[2:, 8] - t[2:, 0]) / 1.9e3).mean():7.2f}")
print("first step:", [round(float(x), 2) for x in d[0]])
