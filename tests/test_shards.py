"""Feature shards: file format round trip on the CPU; the pinned double-buffered feeder on the GPU."""
import numpy as np
import pytest
import torch

from oracle import salstm_oracle as O


def _videos(n, T, Fa, Fv, L, V, seed):
    g = np.random.default_rng(seed)
    au, vi, cp = [], [], []
    for i in range(n):
        t = int(g.integers(1, T + 1))
        au.append(g.integers(0, 256, (t, Fa)).astype(np.float32))
        vi.append((np.maximum(g.standard_normal((t, Fv)), 0) * 10).astype(np.float32))
        k = int(g.integers(3, L + 1))
        cp.append(np.concatenate([[1], g.integers(4, V, k - 2), [2]]).astype(np.int64))
    return au, vi, cp


def test_shard_roundtrip_ragged(tmp_path):
    from salstm.shards import ShardReader, write_shard
    T, Fa, Fv, L, V = 9, 8, 16, 7, 50
    au, vi, cp = _videos(11, T, Fa, Fv, L, V, 0)
    info = write_shard(str(tmp_path / "a.shard"), au, vi, cp, T=T, L=L)
    assert info["N"] == 11 and info["T"] == T and info["L"] == L
    r = ShardReader(str(tmp_path / "a.shard"))
    assert (len(r), r.T, r.Fa, r.Fv, r.L) == (11, T, Fa, Fv, L)
    idx = [7, 2, 10]
    a, v, c, n = r.batch(idx)
    assert a.dtype == torch.bfloat16 and a.shape == (3, T, Fa) and v.shape == (3, T, Fv) and c.shape == (L, 3)
    for j, i in enumerate(idx):
        t = au[i].shape[0]
        assert int(n[j]) == t
        assert torch.equal(a[j, :t], torch.from_numpy(au[i]).bfloat16())          # one round-to-nearest-even, as the bf16 path
        assert torch.equal(v[j, :t], torch.from_numpy(vi[i]).bfloat16())
        assert float(a[j, t:].abs().sum()) == 0 and float(v[j, t:].abs().sum()) == 0     # zero padding (get_loader.py:403-413)
        k = len(cp[i])
        assert c[:k, j].tolist() == cp[i].tolist() and (c[k:, j] == 0).all()       # PAD id 0, time-first
    # a contiguous range takes the slab path and gives the same rows
    a2, v2, c2, n2 = r.batch(range(3, 8))
    a3, v3, c3, n3 = r.batch([3, 4, 5, 6, 7][::-1])
    assert torch.equal(a2, a3.flip(0)) and torch.equal(c2, c3.flip(1)) and torch.equal(n2, n3.flip(0))


def test_shard_truncates_and_rejects_garbage(tmp_path):
    from salstm.shards import ShardReader, write_shard
    au, vi, cp = _videos(4, 12, 8, 8, 10, 30, 1)
    write_shard(str(tmp_path / "b.shard"), au, vi, cp, T=5, L=4)
    r = ShardReader(str(tmp_path / "b.shard"))
    assert r.T == 5 and r.L == 4 and int(r.lengths.max()) <= 5
    (tmp_path / "junk").write_bytes(b"x" * 200)
    with pytest.raises(ValueError, match="not an MVCSHRD1"):
        ShardReader(str(tmp_path / "junk"))


def test_synthetic_batch_through_a_shard_equals_the_loader_contract(tmp_path):
    """A synth_batch written as a shard and read back is the same batch rounded to bf16, with its frame lengths."""
    from salstm.shards import ShardReader, write_shard
    from salstm.synth import frame_lengths
    audio, visual, caps = O.synth_batch(6, 7, 8, 40, Fa=8, Fv=16, seed=3, min_frames=2, min_cap=3)
    write_shard(str(tmp_path / "c.shard"), list(audio), list(visual), list(caps.t()), T=7, L=8)
    a, v, c, n = ShardReader(str(tmp_path / "c.shard")).batch(range(6))
    assert torch.equal(a, audio.bfloat16()) and torch.equal(v, visual.bfloat16()) and torch.equal(c, caps)
    assert torch.equal(n, torch.full((6,), 7, dtype=torch.int32))       # stored length = rows written (padding included)
    assert int(frame_lengths(audio, visual).max()) <= 7


@pytest.mark.gpu
def test_feeder_double_buffering_and_dp_split(tmp_path):
    from salstm.shards import ShardFeeder, ShardReader, write_shard
    dev = torch.device("cuda:0")
    T, Fa, Fv, L, V, N, B = 6, 8, 16, 5, 30, 37, 8
    au, vi, cp = _videos(N, T, Fa, Fv, L, V, 4)
    write_shard(str(tmp_path / "d.shard"), au, vi, cp, T=T, L=L)
    r = ShardReader(str(tmp_path / "d.shard"))
    seen = []
    feeder = ShardFeeder(r, B, dev, shuffle=True, seed=5, drop_last=False)
    assert len(feeder) == 5
    held = []
    for a, v, c, n in feeder:
        assert a.is_cuda and a.dtype == torch.bfloat16 and c.shape[0] == L
        held.append((a.clone(), v.clone(), c.clone(), n.clone()))           # a slot is reused two batches later
        torch.cuda._sleep(2_000_000)                                        # consumer still busy when the next upload starts
    order = np.arange(N); np.random.default_rng(5).shuffle(order)
    for i, (a, v, c, n) in enumerate(held):
        ra, rv, rc, rn = r.batch(order[i * B:(i + 1) * B])
        assert torch.equal(a.cpu(), ra) and torch.equal(v.cpu(), rv) and torch.equal(c.cpu(), rc) and torch.equal(n.cpu(), rn)
    # data-parallel split: ranks see disjoint halves of each global batch
    r0 = [n.shape[0] for *_, n in ShardFeeder(r, 4, dev, rank=0, world=2)]
    b0 = next(iter(ShardFeeder(r, 4, dev, rank=0, world=2)))[0].cpu()
    b1 = next(iter(ShardFeeder(r, 4, dev, rank=1, world=2)))[0].cpu()
    ra = r.batch(range(8))[0]
    assert r0 == [4] * 4 and torch.equal(b0, ra[:4]) and torch.equal(b1, ra[4:])
    # decoding feed: no captions
    a, v, c, n = next(iter(ShardFeeder(r, 8, dev, with_captions=False)))
    assert c is None and a.shape == (8, T, Fa)
    # page-locked shard: consecutive rows upload straight out of the shard (no staging), batches optionally shuffled whole
    rp = ShardReader(str(tmp_path / "d.shard"), pin=True)
    assert rp.pinned and rp.host_tensors()[0].is_pinned()
    got = [(a.cpu(), c.cpu(), n.cpu()) for a, v, c, n in ShardFeeder(rp, B, dev, drop_last=False)]
    for i, (a, c, n) in enumerate(got):
        ra, _, rc, rn = r.batch(range(i * B, min(N, (i + 1) * B)))
        assert torch.equal(a, ra) and torch.equal(c, rc) and torch.equal(n, rn)
    starts = sorted(int(n[0]) for *_, n in [(None, None, x[2]) for x in got])
    fb = ShardFeeder(rp, B, dev, shuffle="batches", seed=3)
    firsts = [a.cpu() for a, *_ in fb]
    ref = {bytes(r.batch(range(lo, lo + B))[0].view(torch.int16).numpy().tobytes()) for lo in range(0, N - B + 1, B)}
    assert len(firsts) == 4 and {bytes(a.view(torch.int16).numpy().tobytes()) for a in firsts} == ref
