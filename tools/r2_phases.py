"""Phase timing of the projected-keys persistent kernels (recur2_fwd.cu / recur2_bwd.cu), CTA 0, SM clock, one train step.
Stamps 0-6 come from thread 0 (row-owner warps), 7-9 from thread 256 (recurrent-GEMM warps)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "multimodal-video-captioning_b200"), ROOT]
import torch
import bench as Bn
from salstm import cabi
import losses as Lm
dev = torch.device("cuda:0")
lib = cabi.lib()
wl = sys.argv[1] if len(sys.argv) > 1 else "train"
shape = Bn.SHAPES["msvd"]
B, T, L, V = shape
model = Bn.build_model(wl, dev, "bf16")
b = [t.to(dev) for t in Bn.make_batches(shape, 1)[0]]
loss_fn = Lm.ModalityWiseReconstructionLossBuilder(rec_type="none", **Bn.LAMBDAS)
S = L - 1
def step():
    out, ar, vr = model(b[0], b[1], b[2])
    loss_fn(out, b[2], b[0], ar, b[1], vr)[0].backward()
for _ in range(3):
    step()
GHZ = 1.9e3
def show(t, names, pairs):
    print("  per-step (us, mean over steps 2..S-2):")
    for n, (a, c) in zip(names, pairs):
        d = (t[2:-1, c] - t[2:-1, a]) / GHZ
        print(f"    {n:44s} {d.mean():7.2f}  (min {d.min():6.2f} max {d.max():6.2f})")
    print(f"    {'step period (stamp 0 -> next stamp 0)':44s} {((t[3:, 0] - t[2:-1, 0]) / GHZ).mean():7.2f}")
buf = torch.zeros(10 * S, dtype=torch.int64, device=dev)
lib.mvc_debug_set_recur_prof(cabi.ptr(buf)); lib.mvc_debug_set_recur_bwd_prof(None)
step(); torch.cuda.synchronize()
lib.mvc_debug_set_recur_prof(None)
t = buf.cpu().view(S, 10).double()
print("recur2_fwd")
show(t, ["wait h rows of the cluster (hb_full)", "wq slice + scatter + wait q_full", "scores + softmax", "P-sum out of TMEM",
         "wait X (recurrent GEMM done)", "cell + publish h", "GEMM group: Y wait + TMA issue (from step start)",
         "GEMM group: MMA + park + part_full", "GEMM group: reduce + gh store + X"],
     [(0, 1), (1, 2), (2, 3), (3, 4), (4, 5), (5, 6), (0, 7), (7, 8), (8, 9)])
buf.zero_()
lib.mvc_debug_set_recur_bwd_prof(cabi.ptr(buf))
step(); torch.cuda.synchronize()
lib.mvc_debug_set_recur_bwd_prof(None)
t = buf.cpu().view(S, 10).double()
print("recur2_bwd")
show(t, ["prefetch + wait X + wait dh_att", "cell backward + publish dG", "dalpha (TMEM) + softmax Jacobian", "dpre + dwq + exchange",
         "wait dq_full", "dh_att mma + scatter", "GEMM group: Y wait + TMA issue (from step start)",
         "GEMM group: MMA + park + part_full", "GEMM group: reduce + ghb store + X"],
     [(0, 1), (1, 2), (2, 3), (3, 4), (4, 5), (5, 6), (0, 7), (7, 8), (8, 9)])
