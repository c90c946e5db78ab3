import os, sys
sys.path[:0] = ["/root/repo/multimodal-video-captioning_b200", "/root/repo"]
import torch
from salstm import cabi
lib = cabi.lib(); dev = torch.device("cuda:0")
def run(M, N, K, reps=30):
    a = torch.randn(M, K, device=dev).bfloat16(); b = torch.randn(N, K, device=dev).bfloat16()
    C = torch.empty(M, N, device=dev)
    st = cabi.stream_ptr()
    f = lambda: cabi.check(lib.mvc_gemm_bf16(M, N, K, cabi.ptr(a), K, cabi.ptr(b), K, 0.0, cabi.ptr(C), N, None, None, 0, st))
    for _ in range(5): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    print(f"M={M} N={N} K={K}: {us:7.1f} us  {2*M*N*K/us/1e6:7.1f} TFLOP/s")
for K in (800, 1600, 3208, 6400): run(2944, 512, K)
for K in (544, 2176, 4352): run(5632, 256, K)
for K in (736, 2944, 5888): run(2048, 512, K)
for K in (304, 1216): run(2944, 2048, K)
for K in (736, 2944): run(3201, 512, K)
