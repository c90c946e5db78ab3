"""Where does the end-to-end train step lose time against the device-resident one?  Variants of bench.py's e2e loop."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "multimodal-video-captioning_b200"), ROOT]
import torch
import bench as Bn
from salstm.trainer import FlatClipAdam
import losses as Lm

dev = torch.device("cuda:0")
w = Bn.WORKLOADS["train"]
shape = Bn.SHAPES[w["shape"]]
B, T, L, V = shape
model = Bn.build_model("train", dev, "bf16")
host = Bn.make_batches(shape, 4)
pinned = [tuple(t.pin_memory() for t in b) for b in host]
resident = [tuple(t.to(dev) for t in b) for b in host]
loss_fn = Lm.ModalityWiseReconstructionLossBuilder(rec_type="none", **Bn.LAMBDAS)
opt = FlatClipAdam(model.parameters(), lr=1e-4)


def step(b):
    opt.zero_grad()
    out, ar, vr = model(b[0], b[1], b[2])
    terms = loss_fn(out, b[2], b[0], ar, b[1], vr)
    terms[0].mean().backward()
    opt.step()
    return terms[0]


copy_stream = torch.cuda.Stream(device=dev)
slots = [tuple(torch.empty_like(t, device=dev) for t in host[0]) for _ in range(2)]
ready = [torch.cuda.Event(), torch.cuda.Event()]
freed = [torch.cuda.Event(), torch.cuda.Event()]
host_res = [torch.empty(1).pin_memory() for _ in range(2)]
res_done = [torch.cuda.Event(), torch.cuda.Event()]


def run(n, copies, readback, sync_every):
    for s in range(2):
        freed[s].record(torch.cuda.current_stream())

    def prefetch(i):
        s = i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(freed[s])
            for dst, src in zip(slots[s], pinned[i % 4]):
                dst.copy_(src, non_blocking=True)
            ready[s].record(copy_stream)
    if copies:
        prefetch(0)
    for i in range(n):
        s = i % 2
        if copies:
            if i + 1 < n:
                prefetch(i + 1)
            torch.cuda.current_stream().wait_event(ready[s])
            r = step(slots[s])
            freed[s].record(torch.cuda.current_stream())
        else:
            r = step(resident[i % 4])
        if readback:
            host_res[s].copy_(r.detach().reshape(1), non_blocking=True)
            res_done[s].record(torch.cuda.current_stream())
            if i >= 1 and sync_every:
                res_done[1 - s].synchronize()
                _ = host_res[1 - s][0].item()
    torch.cuda.synchronize()


for name, kw in [("resident, no readback", dict(copies=False, readback=False, sync_every=False)),
                 ("resident, readback + sync one step behind", dict(copies=False, readback=True, sync_every=True)),
                 ("H2D prefetch, no readback", dict(copies=True, readback=False, sync_every=False)),
                 ("H2D prefetch, readback, no host sync", dict(copies=True, readback=True, sync_every=False)),
                 ("H2D prefetch, readback + sync (bench e2e)", dict(copies=True, readback=True, sync_every=True))]:
    run(5, **kw)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    run(40, **kw)
    e1.record()
    torch.cuda.synchronize()
    print(f"{name:48s} {e0.elapsed_time(e1) / 40:.3f} ms/step (wall {(time.perf_counter() - t0) * 1e3 / 40:.3f})")
