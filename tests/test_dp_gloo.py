"""World-size-2 data-parallel plumbing on CPU (gloo): batch sharding, the flat gradient buffer and its single
all-reduce.  The arithmetic itself is CUDA-only and covered by the -m gpu tests; here each rank
fills its parameter gradients with a known function of (rank, shard) and we check what the all-reduce leaves in
the flat buffer and in the parameter views."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import PKG, ROOT


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    for p in (PKG, ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from salstm.trainer import FlatClipAdam, shard_batch
        torch.manual_seed(0)                       # replicated parameters
        net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Linear(5, 3))
        extra = torch.nn.Parameter(torch.zeros(4))  # never receives a gradient (like AVCaptioningDual.output_fc)
        B = 7                                       # odd batch: ragged shards
        g = torch.Generator().manual_seed(1)
        audio = torch.randn(B, 4, 6, generator=g)
        visual = torch.randn(B, 4, 2, generator=g)
        caps = torch.randint(0, 9, (5, B), generator=g)
        a, v, c = shard_batch(audio, visual, caps, rank, world)
        lo, hi = rank * 4, min(B, rank * 4 + 4)
        assert torch.equal(a, audio[lo:hi]) and torch.equal(v, visual[lo:hi]) and torch.equal(c, caps[:, lo:hi])
        loss = net(a.mean(1)).pow(2).sum()
        loss.backward()
        local = [p.grad.clone() for p in net.parameters()]
        opt = FlatClipAdam(list(net.parameters()) + [extra], lr=1e-3, world_size=world)
        opt.all_reduce_grads()
        # the flat buffer holds exactly the live parameters, in order, each starting on a 16-byte boundary (numel rounded
        # up to a multiple of 4), and the .grad tensors are views of it
        al = lambda k: (k + 3) // 4 * 4
        assert opt.flat_g.numel() == sum(al(p.numel()) for p in net.parameters())
        assert extra.grad is None
        off = 0
        for p in net.parameters():
            assert p.grad.data_ptr() == opt.flat_g[off:off + p.numel()].data_ptr()
            assert p.data.data_ptr() == opt.flat_p[off:off + p.numel()].data_ptr()
            assert p.data.data_ptr() % 16 == 0 and p.grad.data_ptr() % 16 == 0
            off += al(p.numel())
        gathered = [None] * world
        dist.all_gather_object(gathered, [t.tolist() for t in local])
        for i, p in enumerate(net.parameters()):
            want = sum(torch.tensor(gathered[r][i]) for r in range(world))
            torch.testing.assert_close(p.grad, want)
        # a second backward accumulates into the same views; zero_grad clears the flat buffer
        opt.zero_grad()
        assert float(opt.flat_g.abs().sum()) == 0.0
        net(a.mean(1)).pow(2).sum().backward()
        torch.testing.assert_close(net[0].weight.grad, local[0])
        try:
            opt.step()
            ok = False
        except RuntimeError as e:                   # no CPU fallback for the update kernel
            ok = "CUDA" in str(e)
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_dp_world2_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    res = sorted(q.get(timeout=5) for _ in range(world))
    assert res == [(0, True), (1, True)]
