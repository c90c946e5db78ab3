"""Time the gradient all-reduce of the C2 model (9.41 M fp32 = 37.7 MB) alone, eager and inside a CUDA graph.
torchrun --nproc-per-node N tools/allreduce_probe.py [numel]"""
import os, sys, datetime
import torch, torch.distributed as dist
rank, lr, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=120))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 9414573
for dtype in (torch.float32, torch.bfloat16):
    x = torch.randn(n, device=dev).to(dtype)
    for _ in range(10):
        dist.all_reduce(x)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        dist.all_reduce(x)
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / 50], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        us = float(t) * 1e3
        print(f"world {world} {dtype} {n * x.element_size() / 1e6:.1f} MB: {us:.1f} us  algbw {n * x.element_size() / us / 1e3:.0f} GB/s  "
              f"env MIN_NCHANNELS={os.environ.get('NCCL_MIN_NCHANNELS')} ALGO={os.environ.get('NCCL_ALGO')} PROTO={os.environ.get('NCCL_PROTO')}")
dist.destroy_process_group()
