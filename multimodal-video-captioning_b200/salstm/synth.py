"""Synthetic MSVD / MSR-VTT-shaped batches in the loader's layout, for benchmarks and smoke runs.

Product-side generator (bench.py's B200 arm must not import anything from oracle/); tests/test_oracle_golden.py
checks that it produces exactly the batches oracle.salstm_oracle.synth_batch does."""
import torch

PAD, SOS, EOS = 0, 1, 2        # get_loader.py:25-26


def synth_batch(B: int, T: int, L: int, V: int, Fa=128, Fv=2048, seed=1, min_frames=4, min_cap=8):
    """CustomCollateAV layout (get_loader.py:403-413): audio [B,T,Fa] f32 holding integers 0..255 (VGGish
    post-processed range), visual [B,T,Fv] f32 = 10*relu(randn) (max ~48 like Inception pool features), trailing
    frames zero-padded per sample; captions [L,B] i64: SOS first, EOS at len-1, PAD after."""
    g = torch.Generator().manual_seed(seed)
    audio = torch.randint(0, 256, (B, T, Fa), generator=g).float()
    visual = torch.relu(torch.randn(B, T, Fv, generator=g)) * 10.0
    nfr = torch.randint(min(min_frames, T), T + 1, (B,), generator=g)
    keep = (torch.arange(T).unsqueeze(0) < nfr.unsqueeze(1)).unsqueeze(2)
    audio, visual = audio * keep, visual * keep
    lens = torch.randint(min(min_cap, L), L + 1, (B,), generator=g)
    lens[0] = L                                   # pad_sequence guarantees one full-length caption
    caps = torch.randint(4, V, (L, B), generator=g)
    pos = torch.arange(L).unsqueeze(1)
    caps = torch.where(pos == 0, torch.full_like(caps, SOS), caps)
    caps = torch.where(pos == (lens - 1).unsqueeze(0), torch.full_like(caps, EOS), caps)
    caps = torch.where(pos >= lens.unsqueeze(0), torch.full_like(caps, PAD), caps)
    return audio, visual, caps


def frame_lengths(audio, visual):
    """Per-sample count of non-padding frames [B] int32 (the length tensor the reference's collate never passes,
    get_loader.py:403-413): index of the last frame with any non-zero feature, plus one."""
    nz = (audio != 0).any(-1) | (visual != 0).any(-1)
    T = nz.shape[1]
    idx = torch.arange(1, T + 1, device=nz.device).unsqueeze(0) * nz
    return idx.max(1).values.to(torch.int32)
