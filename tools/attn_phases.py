"""Phase timing (%globaltimer, ns) of the staged soft-attention forward kernel over one greedy / beam decode.
Usage: python tools/attn_phases.py [greedy|beam]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "multimodal-video-captioning_b200"), ROOT]
import torch
import bench as Bn
from salstm import cabi
workload = sys.argv[1] if len(sys.argv) > 1 else "greedy"
dev = torch.device("cuda:0")
lib = cabi.lib()
shape = Bn.SHAPES[Bn.WORKLOADS[workload]["shape"]]
B, T, L, V = shape
model = Bn.build_model(workload, dev, "bf16")
b = [t.to(dev) for t in Bn.make_batches(shape, 1)[0]]


def run():
    if workload == "beam":
        from salstm import functional as Fn
        dec = model.decoder
        return Fn.decoder_beam(dec._dims(B, T, L), b[0], b[1], dec._params(), 5, 0.0)
    return model.decoder.greedy_ids((b[0], b[1]), L)


for _ in range(3):
    run()
NC = B * 8
buf = torch.zeros(NC * 8, dtype=torch.int64, device=dev)
lib.mvc_debug_set_attn_prof(cabi.ptr(buf))
run()
torch.cuda.synchronize()
lib.mvc_debug_set_attn_prof(None)
t = buf.cpu().view(NC, 8).double()
live = t[:, 0] > 0
t = t[live]
print(f"{int(live.sum())} CTAs in the last launch; kernel span {(t[:, 7].max() - t[:, 0].min()) / 1e3:.2f} us")
names = ["setup + bulk issue + uk ldg issue", "pdl wait (+ cluster wait)", "wq -> sQ", "scores (+ cluster exchange)",
         "soft-max", "wait keys", "context + store (last query)"]
d = (t[:, 1:8] - t[:, 0:7]) / 1e3
for i, n in enumerate(names):
    print(f"  {n:36s} mean {d[:, i].mean():7.2f}  p10 {d[:, i].quantile(0.1):7.2f}  p90 {d[:, i].quantile(0.9):7.2f} us")
life = (t[:, 7] - t[:, 0]) / 1e3
print(f"  CTA lifetime mean {life.mean():.2f} us; start offsets p10/p50/p90: "
      f"{[round(float(x), 2) for x in ((t[:, 0] - t[:, 0].min()) / 1e3).quantile(torch.tensor([0.1, 0.5, 0.9], dtype=torch.float64))]}")
