"""Kernel timeline of ONE graphed C2 train step (torch.profiler / CUPTI activity records): start, duration and stream of
every kernel relative to the first one, the idle gaps of the union, and which kernels run while nothing else does
(= the critical path).  Numbers under a profiler are not bench values; the shares are what matter.

    python tools/step_timeline.py [train|train_dual|recnet_global|recnet_local] > profiles/step_timeline_r2.txt
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "multimodal-video-captioning_b200"), ROOT]
import torch
from torch.profiler import profile, ProfilerActivity
import bench as Bn
from salstm.trainer import FlatClipAdam, GraphedTrainStep
import losses as Lm

wl = sys.argv[1] if len(sys.argv) > 1 else "train"
# under torchrun (WORLD_SIZE > 1): data-parallel step with the gradient exchange; rank 0 prints its own timeline
world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
    if rank != 0:
        sys.stdout = open(os.devnull, "w")
w = Bn.WORKLOADS[wl]
shape = Bn.SHAPES[w["shape"]]
model = Bn.build_model(wl, dev, "bf16")
bs = [tuple(t.to(dev) for t in b) for b in Bn.make_batches(shape, 4, seed0=1 + 100 * rank)]
loss_fn = Lm.ModalityWiseReconstructionLossBuilder(rec_type=w["rec"], **Bn.LAMBDAS)
opt = FlatClipAdam(model.parameters(), lr=1e-4, world_size=world if world > 1 else None)
step = GraphedTrainStep(model, loss_fn, opt, bs[0], slots=4)
for ins, src in zip(step.input_slots, bs):
    for dst, t_ in zip(ins, src):
        dst.copy_(t_)
bs = step.input_slots                     # batches live in the graphs' own input tensors: no staging copy
for i in range(6):
    step(*bs[i % 4][:3])
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for i in range(3):
        step(*bs[i % 4][:3])
        torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range.end > e.time_range.start]
evs.sort(key=lambda e: e.time_range.start)
# split into the three steps at the largest two gaps
run_end, idle_before = evs[0].time_range.end, {}
for i in range(1, len(evs)):
    idle_before[i] = evs[i].time_range.start - run_end          # GPU idle before kernel i (all streams)
    run_end = max(run_end, evs[i].time_range.end)
cut = sorted(sorted(idle_before, key=idle_before.get, reverse=True)[:2])
steps = [evs[:cut[0]], evs[cut[0]:cut[1]], evs[cut[1]:]]
ks = steps[-1]
t0 = ks[0].time_range.start
print("# %s: %d kernels / memsets in the last profiled step; columns: start us, duration us, stream, exclusive us, name" %
      (wl, len(ks)))
# exclusive time: part of a kernel's interval during which no other kernel is running
pts = sorted(set([e.time_range.start for e in ks] + [e.time_range.end for e in ks]))
excl = {id(e): 0.0 for e in ks}
idle = 0.0
for a, b in zip(pts[:-1], pts[1:]):
    live = [e for e in ks if e.time_range.start <= a and e.time_range.end >= b]
    if len(live) == 1:
        excl[id(live[0])] += b - a
    elif not live:
        idle += b - a
tot = ks[-1].time_range.end - t0
for e in ks:
    stream = getattr(e, "device_index", 0)
    try:
        stream = e.device_resource_id
    except Exception:
        pass
    print("%9.1f %8.1f  s%-3s %8.1f  %s" % (e.time_range.start - t0, e.time_range.end - e.time_range.start, stream,
                                            excl[id(e)], e.name[:110]))
print("# span %.1f us; idle (no kernel running) %.1f us; sum of kernel durations %.1f us" %
      (tot, idle, sum(e.time_range.end - e.time_range.start for e in ks)))
agg = {}
for e in ks:
    n = e.name.split("(")[0][:70]
    a = agg.setdefault(n, [0, 0.0, 0.0])
    a[0] += 1
    a[1] += e.time_range.end - e.time_range.start
    a[2] += excl[id(e)]
print("# by kernel: launches, total us, exclusive us (alone on the GPU)")
for n, a in sorted(agg.items(), key=lambda kv: -kv[1][2]):
    print("#  %3d %9.1f %9.1f  %s" % (a[0], a[1], a[2], n))
if world > 1:
    torch.cuda.synchronize()
    dist.barrier()
    dist.destroy_process_group()
