"""Where does the big tcgen05 GEMM spend its time?  K sweep (mainloop share) and output-format sweep (epilogue share)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "multimodal-video-captioning_b200"), ROOT]
import torch
from salstm import cabi
lib = cabi.lib()
dev = torch.device("cuda:0")
def run(M, N, K, c32, c16, reps=30):
    a = torch.randn(M, K, device=dev).bfloat16(); b = torch.randn(N, K, device=dev).bfloat16()
    C = torch.empty(M, N, device=dev) if c32 else None
    Cb = torch.empty(M, N, device=dev, dtype=torch.bfloat16) if c16 else None
    st = cabi.stream_ptr()
    f = lambda: cabi.check(lib.mvc_gemm_bf16(M, N, K, cabi.ptr(a), K, cabi.ptr(b), K, 0.0, cabi.ptr(C), N, None, cabi.ptr(Cb), N, st))
    for _ in range(5): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    print(f"M={M} N={N} K={K} C32={int(c32)} C16={int(c16)}: {us:7.1f} us  {2*M*N*K/us/1e6:7.1f} TFLOP/s")
for K in (544, 1088, 2176, 4352, 8704):
    run(5632, 2048, K, False, True)
run(5632, 2048, 2176, True, False)
run(5632, 2048, 2176, True, True)
for M in (2368, 4736, 9472, 18944):      # 148 SMs x {1,2,4,8} tiles of 128x256 at N=2048 -> whole waves
    run(M, 2048, 2176, False, True)
