"""Phase timing (%globaltimer) of the one-tile tcgen05 GEMM kernel for launches of shape (M, N, K) inside a workload.
Usage: python tools/gemm_phases.py <workload> M N K"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "multimodal-video-captioning_b200"), ROOT]
import torch
import bench as Bn
from salstm import cabi
from salstm.trainer import FlatClipAdam
import losses as Lm
workload = sys.argv[1]
M, N, K = (int(x) for x in sys.argv[2:5])
dev = torch.device("cuda:0")
lib = cabi.lib()
w = Bn.WORKLOADS[workload]
shape = Bn.SHAPES[w["shape"]]
B, T, L, V = shape
model = Bn.build_model(workload, dev, "bf16")
b = [t.to(dev) for t in Bn.make_batches(shape, 1)[0]]
training = workload in ("train", "recnet_global", "recnet_local")
if training:
    loss_fn = Lm.ModalityWiseReconstructionLossBuilder(rec_type=w["rec"], **Bn.LAMBDAS)
    opt = FlatClipAdam(model.parameters(), lr=1e-4)


def run():
    if workload == "beam":
        from salstm import functional as Fn
        dec = model.decoder
        return Fn.decoder_beam(dec._dims(B, T, L), b[0], b[1], dec._params(), 5, 0.0)
    if not training:
        return model.decoder.greedy_ids((b[0], b[1]), L)
    opt.zero_grad()
    out, ar, vr = model(b[0], b[1], b[2])
    terms = loss_fn(out, b[2], b[0], ar, b[1], vr)
    terms[0].mean().backward()
    opt.step()


for _ in range(3):
    run()
NC = 4096
buf = torch.zeros(NC * 8, dtype=torch.int64, device=dev)
lib.mvc_debug_set_gemm_prof(cabi.ptr(buf), M, N, K)
run()
torch.cuda.synchronize()
lib.mvc_debug_set_gemm_prof(None, 0, 0, 0)
t = buf.cpu().view(NC, 8).double()
live = t[:, 0] > 0
t = t[live]
if not len(t):
    sys.exit("no launch of that shape went through the one-tile kernel")
split = bool((t[:, 5] > 0).all())
print(f"{len(t)} CTAs in the last launch ({'split-K' if split else 'no split'}); kernel span {(t[:, 7].max() - t[:, 0].min()) / 1e3:.2f} us")
pairs = [("setup (barriers, TMEM alloc)", 0, 1), ("producer: wait upstream kernel", 1, 2), ("producer: issue all TMA loads", 2, 3),
         ("entry -> accumulator complete", 0, 4)]
if split:
    pairs += [("park partial + cluster barrier", 4, 5), ("DSMEM reduce + epilogue stores", 5, 6), ("final cluster barrier + exit", 6, 7)]
else:
    pairs += [("epilogue + exit", 4, 7)]
for n, i, j in pairs:
    d = (t[:, j] - t[:, i]) / 1e3
    print(f"  {n:36s} mean {d.mean():7.2f}  p10 {d.quantile(0.1):7.2f}  p90 {d.quantile(0.9):7.2f} us")
print(f"  CTA lifetime mean {((t[:, 7] - t[:, 0]) / 1e3).mean():.2f} us; start spread {(t[:, 0].max() - t[:, 0].min()) / 1e3:.2f} us")
