"""In-situ (warm, inside the real train step) per-kernel-class timings via mvc_prof_arm, plus the host-side
cost of issuing one step.  Usage: python tools/kernel_times.py [steps]"""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "multimodal-video-captioning_b200"), ROOT]
import torch
import bench as Bn
from salstm import cabi
from salstm.trainer import FlatClipAdam
import losses as Lm

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
workload = sys.argv[2] if len(sys.argv) > 2 else "train"
dev = torch.device("cuda:0")
lib = cabi.lib()
w = Bn.WORKLOADS[workload]
shape = Bn.SHAPES[w["shape"]]
B, T, L, V = shape
model = Bn.build_model(workload, dev, "bf16")
batches = [tuple(t.to(dev) for t in b) for b in Bn.make_batches(shape, 2)]
loss_fn = Lm.ModalityWiseReconstructionLossBuilder(rec_type=w["rec"], **Bn.LAMBDAS)
opt = FlatClipAdam(model.parameters(), lr=1e-4)
training = workload.startswith("train") or workload.startswith("recnet")


def step(b):
    if workload == "beam":
        from salstm import functional as Fn
        dec = model.decoder
        return Fn.decoder_beam(dec._dims(b[0].shape[0], b[0].shape[1], L), b[0], b[1], dec._params(), 5, 0.0)
    if not training:
        return model.decoder.greedy_ids((b[0], b[1]), L)
    opt.zero_grad()
    out, ar, vr = model(b[0], b[1], b[2])
    terms = loss_fn(out, b[2], b[0], ar, b[1], vr)
    terms[0].mean().backward()
    opt.step()
    return terms[0]


for i in range(5):
    step(batches[i % 2])
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(steps):
    step(batches[i % 2])
t_issue = time.perf_counter() - t0
torch.cuda.synchronize()
t_all = time.perf_counter() - t0
print(f"step: {t_all / steps * 1e3:.3f} ms wall; host issue time {t_issue / steps * 1e3:.3f} ms/step")

S = L - 1
classes = [
    ("gemm_tc ALL", (1, -1, -1, -1)),
    ("gemm_tc gates+cell (B,4H,F+H)", (1, B, 2048, 2688)),
    ("gemm_tc gates K=2992 (greedy)", (1, B, 2048, 2992)),
    ("gemm_tc wq (B,A,H)", (1, B, 256, 512)),
    ("gemm_tc dxh (B,F+H,4H)", (1, B, 2688, 2048)),
    ("gemm_tc dh+=dwq.W (B,H,A)", (1, B, 512, 256)),
    ("gemm_tc vocab fwd", (1, S * B, V, 512)),
    ("gemm_tc vocab step", (1, B, V, 512)),
    ("gemm_tc embtab", (1, V, 2048, 304)),
    ("beam gates+cell (5B,4H,F+H)", (1, 5 * B, 2048, 2688)),
    ("beam vocab topk (5B,V+aux,H)", (1, 5 * B, (V + 255) // 256 * 256 + 256, 512)),
    ("greedy vocab argmax (B,V+aux,H)", (1, B, (V + 255) // 256 * 256 + 256, 512)),
    ("gemm_tc gates greedy (B,4H,F+H)", (1, B, 2048, 2688)),
    ("gemm_tc uk", (1, B * T, 256, 2176)),
    ("gemm_tc P = keys.Wc^T (BT,4H,F)", (1, B * T, 2048, 2176)),
    ("gemm_tn dW_ih|dW_hh merged (4H,E+F+H,SB)", (1, 2048, 300 + 2176 + 512, S * B)),
    ("gemm_tn dW_c (4H,F,SB)", (1, 2048, 2176, S * B)),
    ("gemm_tn dW_hh (4H,H,SB)", (1, 2048, 512, S * B)),
    ("gemm_tc gx (SB,4H,Ep)", (1, S * B, 2048, 304)),
    ("gemm_tn dW_cat (4H,F+H,SB)", (1, 2048, 2688, S * B)),
    ("gemm_tn dW_out (V,H,SB)", (1, V, 512, S * B)),
    ("gemm_nn dhall (SB,H,V)", (1, S * B, 512, V)),
    ("gemm_tn dW_att (A,H,SB)", (1, 256, 512, S * B)),
    ("gemm_tn dU (A,F,BT)", (1, 256, 2176, B * T)),
    ("gemm_tn dW_ie (4H,E,SB)", (1, 2048, 300, S * B)),
    ("gemm_nn dxemb (SB,E,4H)", (1, S * B, 300, 2048)),
    ("recon-g xproj (SB,4F,H)", (1, S * B, 8704, 512)),
    ("recon-g gates+cell (B,4F,F)", (1, B, 8704, 2176)),
    ("recon-g dh (B,F,4F)", (1, B, 2176, 8704)),
    ("recon-g dW_hh (4F,F,SB)", (1, 8704, 2176, S * B)),
    ("recon-g dW_ih (4F,H,SB)", (1, 8704, 512, S * B)),
    ("recon-g dx (SB,H,4F)", (1, S * B, 512, 8704)),
    ("recon-l wq (B,A,F)", (1, B, 256, 2176)),
    ("recon-l gates+cell (B,4F,H+F)", (1, B, 8704, 2688)),
    ("recon-l dxh (B,H+F,4F)", (1, B, 2688, 8704)),
    ("recon-l dq (B,F,A)", (1, B, 2176, 256)),
    ("recon-l dW_hh (4F,F,TB)", (1, 8704, 2176, T * B)),
    ("recon-l dW_ih (4F,H,TB)", (1, 8704, 512, T * B)),
    ("recon-l uk (BL,A,H)", (1, B * L, 256, 512)),
    ("persistent recurrence fwd", (8, -1, -1, -1)),
    ("attention fwd", (3, -1, -1, -1)),
    ("attention bwd", (4, -1, -1, -1)),
    ("cell fwd", (5, -1, -1, -1)),
    ("cell bwd", (6, -1, -1, -1)),
    ("log_softmax rows", (7, -1, -1, -1)),
    ("caption loss (entropy)", (9, -1, -1, -1)),
    ("clip+adam", (10, -1, -1, -1)),
]
tot_known = 0.0
for name, (kid, m, n, k) in classes:
    lib.mvc_prof_arm(kid, m, n, k)
    for i in range(steps):
        step(batches[i % 2])
    torch.cuda.synchronize()
    tot, cnt = C.c_double(0), C.c_longlong(0)
    lib.mvc_prof_collect(C.byref(tot), C.byref(cnt))
    if cnt.value:
        print(f"{name:36s} {cnt.value / steps:7.1f} launches/step  avg {tot.value / cnt.value * 1e3:8.2f} us   "
              f"{tot.value / steps:7.3f} ms/step")
