"""Pre-packed feature shards + pinned double-buffered feeder.

The reference's loader (src/get_loader.py:242-268, 333-343, 403-413) does, per item, two ``np.load`` calls of fp32
``.npy`` feature files, a spaCy tokenisation of the caption and a ``pad_sequence`` -- a few hundred samples/s of
Python, against a decoder that consumes ~10^5 samples/s per GPU.  This module is the B200-side replacement of that
data format:

* ``write_shard``  : offline, once -- packs N videos into ONE file: bf16 audio / visual features zero-padded to a common
                     frame count T (bit-identical to what the bf16 compute path makes of fp32 features: one
                     round-to-nearest-even, see tests/test_gpu_parity.py::test_bf16_feature_shards_are_bit_identical),
                     numericalised captions (int64, PAD-padded to L), and the per-video FRAME-LENGTH tensor the
                     reference's collate never passes on (get_loader.py:403-413).
* ``ShardReader``  : ``np.memmap`` view of a shard (``batch(idx)`` gathers rows without touching the rest of the file) or,
                     with ``pin=True``, the whole shard read once into PAGE-LOCKED host memory (the precomputed features of
                     MSVD are 0.4 GB, of MSR-VTT 1.3 GB): batches of consecutive videos are then uploaded straight out
                     of the shard, with no staging copy on the host.
* ``ShardFeeder``  : iterator of device batches in the model's input contract -- audio [B,T,Fa] bf16, visual [B,T,Fv]
                     bf16, captions [L,B] int64 (time-first, like ``CustomCollateAV``), lengths [B] int32 -- through two
                     device slots: a copy stream uploads batch i+1 while the compute stream runs batch i.  Consecutive
                     rows of a pinned shard go DMA-direct; shuffled rows (or an unpinned shard) are first gathered into
                     pinned staging buffers by a worker thread (a host memcpy of ~25 MB per MSVD batch costs 3-5 ms of one
                     core -- more than the 1.3 ms train step -- so shuffle at shard-writing time, or per epoch over
                     whole batches with ``shuffle="batches"``).

File layout (little endian, every array 64-byte aligned):
    0   : magic  b"MVCSHRD1"
    8   : u32 N, T, Fa, Fv, L, reserved[3]
    64  : i32 lengths[N] | i64 captions[N, L] | bf16 audio[N, T, Fa] | bf16 visual[N, T, Fv]
"""
from __future__ import annotations

import os
import queue
import struct
import threading
from typing import Iterable, Iterator, List, Optional, Sequence, Tuple

import numpy as np
import torch

MAGIC = b"MVCSHRD1"
PAD = 0


def _align(n: int, a: int = 64) -> int:
    return (n + a - 1) // a * a


def _layout(N, T, Fa, Fv, L):
    off = 64
    o_len = off; off = _align(off + 4 * N)
    o_cap = off; off = _align(off + 8 * N * L)
    o_aud = off; off = _align(off + 2 * N * T * Fa)
    o_vis = off; off = _align(off + 2 * N * T * Fv)
    return o_len, o_cap, o_aud, o_vis, off


def _to_bf16_bits(x: torch.Tensor) -> np.ndarray:
    """fp32 -> bf16 (round to nearest even) as uint16 bit patterns."""
    return x.to(torch.float32).contiguous().bfloat16().view(torch.int16).numpy().view(np.uint16)


def write_shard(path: str, audio: Sequence, visual: Sequence, captions: Sequence, T: Optional[int] = None,
                L: Optional[int] = None) -> dict:
    """Pack N videos.  audio[i] : [t_i, Fa], visual[i] : [t_i, Fv] (numpy or torch, any float dtype; what the reference
    keeps as one .npy per video), captions[i] : 1-D int sequence *including* <SOS>/<EOS> (get_loader.py:238-240).
    Frames beyond T / tokens beyond L are truncated; shorter ones zero / PAD padded."""
    N = len(audio)
    assert N == len(visual) == len(captions) and N > 0
    au = [torch.as_tensor(np.asarray(a)) for a in audio]
    vi = [torch.as_tensor(np.asarray(v)) for v in visual]
    cp = [torch.as_tensor(np.asarray(c), dtype=torch.int64).reshape(-1) for c in captions]
    Fa, Fv = au[0].shape[-1], vi[0].shape[-1]
    T = int(T or max(max(a.shape[0], v.shape[0]) for a, v in zip(au, vi)))
    L = int(L or max(c.numel() for c in cp))
    o_len, o_cap, o_aud, o_vis, total = _layout(N, T, Fa, Fv, L)
    tmp = path + ".tmp"
    with open(tmp, "wb") as fh:
        fh.truncate(total)
    mm = np.memmap(tmp, dtype=np.uint8, mode="r+")
    mm[:8] = np.frombuffer(MAGIC, dtype=np.uint8)
    mm[8:40] = np.frombuffer(struct.pack("<8I", N, T, Fa, Fv, L, 0, 0, 0), dtype=np.uint8)
    lens = mm[o_len:o_len + 4 * N].view(np.int32)
    caps = mm[o_cap:o_cap + 8 * N * L].view(np.int64).reshape(N, L)
    aud = mm[o_aud:o_aud + 2 * N * T * Fa].view(np.uint16).reshape(N, T, Fa)
    vis = mm[o_vis:o_vis + 2 * N * T * Fv].view(np.uint16).reshape(N, T, Fv)
    caps[:] = PAD
    for i in range(N):
        ta, tv = min(T, au[i].shape[0]), min(T, vi[i].shape[0])
        aud[i, :ta] = _to_bf16_bits(au[i][:ta])
        vis[i, :tv] = _to_bf16_bits(vi[i][:tv])
        lens[i] = max(ta, tv)
        n = min(L, cp[i].numel())
        caps[i, :n] = cp[i][:n].numpy()
    mm.flush()
    del mm
    os.replace(tmp, path)
    return dict(N=N, T=T, Fa=Fa, Fv=Fv, L=L, bytes=total)


class ShardReader:
    def __init__(self, path: str, pin: bool = False):
        self.path = path
        self.pinned = False
        self._mm = np.memmap(path, dtype=np.uint8, mode="r")
        if self._mm.size < 64 or bytes(self._mm[:8]) != MAGIC:
            raise ValueError(f"{path}: not an MVCSHRD1 feature shard")
        if pin:
            buf = torch.empty(self._mm.size, dtype=torch.uint8).pin_memory()
            with open(path, "rb") as fh:
                fh.readinto(buf.numpy())
            self._pin = buf                       # keeps the page-locked allocation alive
            self._mm = buf.numpy()
            self.pinned = True
        self.N, self.T, self.Fa, self.Fv, self.L = struct.unpack("<8I", bytes(self._mm[8:40]))[:5]
        o_len, o_cap, o_aud, o_vis, total = _layout(self.N, self.T, self.Fa, self.Fv, self.L)
        if self._mm.size < total:
            raise ValueError(f"{path}: truncated shard ({self._mm.size} < {total} bytes)")
        self.lengths = self._mm[o_len:o_len + 4 * self.N].view(np.int32)
        self.captions = self._mm[o_cap:o_cap + 8 * self.N * self.L].view(np.int64).reshape(self.N, self.L)
        self.audio = self._mm[o_aud:o_aud + 2 * self.N * self.T * self.Fa].view(np.uint16).reshape(self.N, self.T, self.Fa)
        self.visual = self._mm[o_vis:o_vis + 2 * self.N * self.T * self.Fv].view(np.uint16).reshape(self.N, self.T, self.Fv)

    def __len__(self):
        return self.N

    def host_tensors(self):
        """The shard's arrays as torch tensors over the same memory (page-locked when the reader was opened with pin=True):
        audio / visual as int16 bit patterns of bf16, captions [N, L] int64, lengths [N] int32."""
        t = lambda a: torch.from_numpy(a.view(np.int16) if a.dtype == np.uint16 else a)
        return t(self.audio), t(self.visual), t(self.captions), t(self.lengths)

    def gather_into(self, idx: np.ndarray, audio_out: np.ndarray, visual_out: np.ndarray, caps_out: np.ndarray,
                    lens_out: np.ndarray) -> None:
        """Rows idx -> preallocated (pinned) staging arrays: audio [B,T,Fa] u16, visual [B,T,Fv] u16, captions [L,B] i64
        (time-first), lengths [B] i32.  Contiguous index ranges copy as one slab."""
        B = len(idx)
        if B and np.all(np.diff(idx) == 1):
            lo = int(idx[0])
            np.copyto(audio_out[:B], self.audio[lo:lo + B])
            np.copyto(visual_out[:B], self.visual[lo:lo + B])
            np.copyto(caps_out[:, :B], self.captions[lo:lo + B].T)
            np.copyto(lens_out[:B], self.lengths[lo:lo + B])
        else:
            np.take(self.audio, idx, axis=0, out=audio_out[:B])
            np.take(self.visual, idx, axis=0, out=visual_out[:B])
            caps_out[:, :B] = self.captions[idx].T
            lens_out[:B] = self.lengths[idx]

    def batch(self, idx) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
        """Host tensors (audio bf16 [B,T,Fa], visual bf16 [B,T,Fv], captions [L,B] i64, lengths [B] i32) for rows idx."""
        idx = np.asarray(idx, dtype=np.int64)
        B = len(idx)
        a = np.empty((B, self.T, self.Fa), np.uint16); v = np.empty((B, self.T, self.Fv), np.uint16)
        c = np.empty((self.L, B), np.int64); n = np.empty((B,), np.int32)
        self.gather_into(idx, a, v, c, n)
        bf = lambda x: torch.from_numpy(x.view(np.int16)).view(torch.bfloat16)
        return bf(a), bf(v), torch.from_numpy(c), torch.from_numpy(n)


class ShardFeeder:
    """Double-buffered device batches from one shard.

        feeder = ShardFeeder(ShardReader(path), batch_size=128, device="cuda:0", shuffle=True, seed=0)
        for audio, visual, captions, lengths in feeder:           # one epoch
            out, ar, vr = model(audio, visual, captions)

    A yielded batch lives in one of two device slots; it stays valid until the batch after next is requested (the
    feeder records, on the consumer's stream, when a slot may be overwritten).  ``rank`` / ``world`` give every
    data-parallel rank its contiguous share of each global batch.  ``with_captions=False`` feeds decoding.
    ``device_slots`` = two (audio, visual, captions) tuples of device tensors to upload into instead of the feeder's own
    (``GraphedTrainStep.input_slots``: the batch lands in the graph's static inputs, no device-to-device copy).
    """

    def __init__(self, reader: ShardReader, batch_size: int, device, shuffle: bool = False, seed: int = 0,
                 drop_last: bool = True, rank: int = 0, world: int = 1, with_captions: bool = True, epochs: int = 1,
                 device_slots=None):
        self.r, self.B, self.dev = reader, int(batch_size), torch.device(device)
        if self.dev.type != "cuda":
            raise RuntimeError("ShardFeeder feeds CUDA devices (the decoder has no CPU path)")
        self.shuffle, self.seed, self.drop_last = shuffle, seed, drop_last
        self.rank, self.world, self.with_captions, self.epochs = rank, world, with_captions, epochs
        T, Fa, Fv, L, B = reader.T, reader.Fa, reader.Fv, reader.L, self.B
        pin = lambda *shape, dtype: torch.empty(*shape, dtype=dtype).pin_memory()
        self._host = [(pin(B, T, Fa, dtype=torch.int16), pin(B, T, Fv, dtype=torch.int16), pin(L, B, dtype=torch.int64),
                       pin(B, dtype=torch.int32)) for _ in range(2)]
        self._np = [tuple(t.numpy().view(np.uint16) if t.dtype == torch.int16 else t.numpy() for t in h) for h in self._host]
        self._devt = [(torch.empty(B, T, Fa, dtype=torch.bfloat16, device=self.dev),
                       torch.empty(B, T, Fv, dtype=torch.bfloat16, device=self.dev),
                       torch.empty(L, B, dtype=torch.int64, device=self.dev),
                       torch.empty(B, dtype=torch.int32, device=self.dev)) for _ in range(2)]
        if device_slots is not None:
            if len(device_slots) < 2:
                raise ValueError("ShardFeeder(device_slots=...): two slots are needed (double buffering)")
            for k in range(2):
                a, v, c = device_slots[k][:3]
                want = self._devt[k]
                for got, ref, name in ((a, want[0], "audio"), (v, want[1], "visual"), (c, want[2], "captions")):
                    if got.shape != ref.shape or got.dtype != ref.dtype or got.device != ref.device or not got.is_contiguous():
                        raise ValueError(f"ShardFeeder(device_slots=...): {name} slot must be a contiguous "
                                         f"{tuple(ref.shape)} {ref.dtype} tensor on {ref.device}")
                self._devt[k] = (a, v, c, want[3])
        self._direct = reader.pinned and shuffle in (False, "batches")
        self._src = reader.host_tensors() if self._direct else None
        self._copy = torch.cuda.Stream(device=self.dev)
        self._ready = [torch.cuda.Event() for _ in range(2)]
        self._freed = [torch.cuda.Event() for _ in range(2)]
        self._uploaded = [torch.cuda.Event() for _ in range(2)]
        self.h2d_bytes_per_batch = 2 * B * T * (Fa + Fv) + 4 * B + (8 * L * B if with_captions else 0)

    def _index_batches(self) -> List[np.ndarray]:
        out = []
        for ep in range(self.epochs):
            order = np.arange(self.r.N, dtype=np.int64)
            if self.shuffle is True:
                np.random.default_rng(self.seed + ep).shuffle(order)
            G = self.B * self.world
            starts = list(range(0, self.r.N, G))
            if self.shuffle == "batches":                      # whole global batches in random order: rows stay consecutive
                np.random.default_rng(self.seed + ep).shuffle(starts)
            for lo in starts:
                g = order[lo:lo + G]
                if len(g) < G and self.drop_last:
                    continue
                per = (len(g) + self.world - 1) // self.world
                mine = g[self.rank * per:(self.rank + 1) * per]
                if len(mine):
                    out.append(mine)
        return out

    def __len__(self):
        return len(self._index_batches())

    def __iter__(self) -> Iterator[Tuple[torch.Tensor, torch.Tensor, Optional[torch.Tensor], torch.Tensor]]:
        batches = self._index_batches()
        staged: "queue.Queue" = queue.Queue(maxsize=1)
        slot_free = [threading.Semaphore(1), threading.Semaphore(1)]      # host staging buffer reusable

        def stage():                          # worker thread: memmap -> pinned staging (numpy copies release the GIL)
            try:
                for i, idx in enumerate(batches):
                    s = i % 2
                    slot_free[s].acquire()
                    if self._direct:          # features go DMA-direct out of the pinned shard: stage the small tensors only
                        lo, n = int(idx[0]), len(idx)
                        if self.with_captions:
                            np.copyto(self._np[s][2][:, :n], self.r.captions[lo:lo + n].T)
                    else:
                        self.r.gather_into(idx, *self._np[s])
                    staged.put((i, len(idx)))
                staged.put(None)
            except BaseException as e:        # surface worker failures in the consumer
                staged.put(e)

        th = threading.Thread(target=stage, daemon=True)
        th.start()
        cur = torch.cuda.current_stream(self.dev)
        for s in range(2):
            self._freed[s].record(cur)

        def upload(item):
            i, n = item
            s = i % 2
            with torch.cuda.stream(self._copy):
                self._copy.wait_event(self._freed[s])           # the consumer is done with the device slot
                hs, ds = self._host[s], self._devt[s]
                if self._direct:
                    lo = int(batches[i][0])
                    ds[0][:n].view(torch.int16).copy_(self._src[0][lo:lo + n], non_blocking=True)
                    ds[1][:n].view(torch.int16).copy_(self._src[1][lo:lo + n], non_blocking=True)
                    ds[3][:n].copy_(self._src[3][lo:lo + n], non_blocking=True)
                else:
                    ds[0].view(torch.int16).copy_(hs[0], non_blocking=True)
                    ds[1].view(torch.int16).copy_(hs[1], non_blocking=True)
                    ds[3].copy_(hs[3], non_blocking=True)
                if self.with_captions:
                    ds[2].copy_(hs[2], non_blocking=True)
                self._uploaded[s].record(self._copy)
                self._ready[s].record(self._copy)
            return s, n

        def release_host(s):                  # staging buffer s may be refilled once its upload has left the host
            self._uploaded[s].synchronize()
            slot_free[s].release()

        nxt = staged.get()
        if isinstance(nxt, BaseException):
            raise nxt
        pending = upload(nxt) if nxt is not None else None
        prev_slot = None
        while pending is not None:
            s, n = pending
            item = staged.get()               # batch i+1 is staged (or the epoch is over)
            if isinstance(item, BaseException):
                raise item
            nxt_pending = upload(item) if item is not None else None
            cur = torch.cuda.current_stream(self.dev)
            cur.wait_event(self._ready[s])
            a, v, c, ln = self._devt[s]
            if n < self.B:
                a, v, c, ln = a[:n], v[:n], c[:, :n], ln[:n]
            yield a, v, (c if self.with_captions else None), ln
            # the consumer asked for the next batch: everything it enqueued on its stream so far used slot s
            self._freed[s].record(torch.cuda.current_stream(self.dev))
            release_host(s)                   # (the upload of batch i finished long ago: no stall)
            pending = nxt_pending
        th.join()
