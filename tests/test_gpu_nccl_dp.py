"""2-rank NCCL data-parallel parity: each rank runs the CUDA path on its shard of the batch,
FlatClipAdam.all_reduce_grads() exchanges the flat gradient buffer over NCCL, and (flat gradient / world) must equal
the MEAN OF THE PER-SHARD oracle gradients (each shard's loss is its own token-mean CE + batch-axis entropy: standard
DDP mean-of-means, the exactness caveat of §8e).  After the step both ranks hold identical parameters.
Needs 2 GPUs: skipped on a single-GPU box (run with `gpurun --gpus 2`)."""
import os
import socket
import sys

import pytest
import torch

from conftest import PKG, ROOT

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, precision, q):
    for p in (PKG, ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from oracle import salstm_oracle as O
        from models import AVCaptioning
        from salstm.trainer import FlatClipAdam, shard_batch
        import losses as Lm

        class Vocab:
            stoi = {"<SOS>": 1, "<EOS>": 2}

            def __len__(self):
                return 97

        V, B, T, L = 97, 10, 6, 7
        gen = torch.Generator().manual_seed(0)
        p = O.init_decoder_params("decoder.", 2176, V, gen=gen)
        p.update(O.init_recon_params("reconstructor.", "global", 512, 2176, gen=gen))
        audio, visual, caps = O.synth_batch(B, T, L, V, seed=5, min_frames=2, min_cap=3)
        audio, visual = audio / 255.0, visual / 10.0
        lam = dict(reg_lambda=0.0005, audio_recon_lambda=0.00005, visual_recon_lambda=0.5)
        model = AVCaptioning(Vocab(), 1.0, "global", device=dev, precision=precision).to(dev)
        sd = model.state_dict(); sd.update({k: p[k].clone() for k in sd}); model.load_state_dict(sd)
        opt = FlatClipAdam(model.parameters(), lr=1e-3, weight_decay=1e-5, clip_value=5.0, world_size=world,
                           fused_comm=False)                 # the NCCL all-reduce path (the fused path is tested below)
        a, v, c = shard_batch(audio, visual, caps, rank, world)
        for it in range(2):                                  # second iteration exercises the gradient arena path
            opt.zero_grad()
            out, ar, vr = model(a.to(dev), v.to(dev), c.to(dev))
            terms = Lm.ModalityWiseReconstructionLoss(out, c.to(dev), a.to(dev), ar, v.to(dev), vr, rec_type="global", **lam)
            terms[0].mean().backward()
            opt.all_reduce_grads()
            if it == 0:
                got = {k: (t.grad / world).detach().cpu() for k, t in model.named_parameters()}
                opt.step()
        # oracle: per-shard gradients on the CPU, averaged
        want = None
        for r in range(world):
            pa = {k: t.clone().double().requires_grad_() for k, t in p.items()}
            sa, sv, sc = shard_batch(audio, visual, caps, r, world)
            o, oar, ovr = O.av_forward(pa, sa.double(), sv.double(), sc, 1.0, "global", hoist=True)
            t = O.modality_wise_loss(o, sc, sa.double(), oar, sv.double(), ovr, rec_type="global", **lam)
            t[0].backward()
            g = {k: x.grad / world for k, x in pa.items()}
            want = g if want is None else {k: want[k] + g[k] for k in g}
        bad = []
        for k in got:
            ref = want[k].float()
            if precision == "fp32":
                if not torch.allclose(got[k], ref, rtol=2e-3, atol=2e-3 * float(ref.abs().max()) + 1e-9):
                    bad.append(f"{k}: max err {float((got[k] - ref).abs().max()):.3e} (ref max {float(ref.abs().max()):.3e})")
            else:
                cs = float((got[k].double().flatten() @ want[k].flatten()) / (got[k].double().norm() * want[k].norm() + 1e-30))
                if cs < 0.999:
                    bad.append(f"{k}: cosine {cs:.5f}")
        # replicas stay identical after the update
        flat = opt.flat_p.clone()
        other = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(other, flat)
        same = all(torch.equal(other[0], o) for o in other)
        # ---- the fused exchange + update over NVSwitch multicast (mvc_clip_adam_multimem) must land on the same parameters
        # as NCCL all-reduce + full-size update, step after step, and leave identical replicas
        fused = "unsupported"
        torch.manual_seed(1)
        twins = []
        for fc, reduce in ((False, None), (True, "peer"), (True, "switch")):
            m2 = AVCaptioning(Vocab(), 1.0, "global", device=dev, precision=precision).to(dev)
            m2.load_state_dict({k: p[k].clone() for k in m2.state_dict()})
            try:
                o2 = FlatClipAdam(m2.parameters(), lr=1e-3, weight_decay=1e-5, clip_value=5.0, world_size=world, fused_comm=fc)
                if reduce:
                    o2._reduce = reduce          # reduce-scatter by peer loads / by multimem.ld_reduce
                for it in range(3):
                    o2.zero_grad()
                    out, ar, vr = m2(a.to(dev), v.to(dev), c.to(dev))
                    Lm.ModalityWiseReconstructionLoss(out, c.to(dev), a.to(dev), ar, v.to(dev), vr, rec_type="global",
                                                      **lam)[0].mean().backward()
                    o2.all_reduce_grads()
                    o2.step()
                twins.append((m2, o2))
            except RuntimeError as e:
                if "multicast" in str(e) or "symmetric" in str(e):
                    break
                raise
        if len(twins) == 3:
            (ma, oa) = twins[0]
            fused = "ok"
            for mb, ob in twins[1:]:
                assert ob._mc is not None
                for (k, pa), pb in zip(ma.named_parameters(), mb.parameters()):
                    if not torch.allclose(pa, pb, rtol=2e-5, atol=2e-6):
                        fused = f"{ob._reduce} {k}: max diff {float((pa - pb).abs().max()):.3e}"
                        break
            flat2 = ob.flat_p[:ob._n_flat].clone()
            other2 = [torch.empty_like(flat2) for _ in range(world)]
            dist.all_gather(other2, flat2)
            if not all(torch.equal(other2[0], o) for o in other2):
                fused = "replicas diverged under the fused exchange"
        q.put((rank, bad, same, fused))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_nccl_dp_flat_gradient_is_mean_of_shard_gradients(precision):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import __graft_entry__ as g
    g.build()
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, precision, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(600)
        assert p.exitcode == 0
    res = sorted(q.get(timeout=5) for _ in range(world))
    for rank, bad, same, fused in res:
        assert not bad, f"rank {rank}:\n" + "\n".join(bad)
        assert same, "replicas diverged after the optimiser step"
        assert fused in ("ok", "unsupported"), f"rank {rank}: fused multicast exchange: {fused}"
    print("fused multicast exchange:", res[0][3])
