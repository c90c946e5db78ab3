"""bf16 vs fp32 agreement of greedy and width-5 beam captions at raw and O(1) feature scale (diagnostic for
tests/test_gpu_parity.py::test_beam5_bf16_agrees_with_fp32_path)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "multimodal-video-captioning_b200"), ROOT, os.path.join(ROOT, "tests")]
import torch
import __graft_entry__ as g
g.build()
from oracle import salstm_oracle as O
from models import AVCaptioning
class Vocab:
    def __init__(self, n): self.n = n; self.stoi = {"<PAD>": 0, "<SOS>": 1, "<EOS>": 2, "<UNK>": 3}
    def __len__(self): return self.n
    def decode_indexes(self, idx): return " ".join(str(int(i)) for i in idx)
def _prefix(ids):
    ids = [int(x) for x in ids]
    return ids[: ids.index(O.EOS) + 1] if O.EOS in ids[1:] else ids
dev = torch.device("cuda:0")
B, T, V = 64, 30, 10547
for scale_name, sa, sv in [("raw", 1.0, 1.0), ("unit", 1 / 255.0, 1 / 48.0)]:
    torch.manual_seed(3)
    m32 = AVCaptioning(Vocab(V), 0.0, "none", device=dev).to(dev)
    with torch.no_grad():
        m32.decoder.out.weight.mul_(8.0)
    mbf = AVCaptioning(Vocab(V), 0.0, "none", device=dev, precision="bf16").to(dev)
    mbf.load_state_dict(m32.state_dict())
    audio, visual, _ = (t.to(dev) for t in O.synth_batch(B, T, 20, V, seed=11, min_frames=10))
    audio = audio * sa; visual = visual * sv
    with torch.no_grad():
        g32 = m32.predict_ids(audio, visual, 20, mode="direct"); gbf = mbf.predict_ids(audio, visual, 20, mode="direct")
        a = m32.predict_ids(audio, visual, 20, mode="beam", beam_width=5)
        b = mbf.predict_ids(audio, visual, 20, mode="beam", beam_width=5)
        b1 = mbf.predict_ids(audio, visual, 20, mode="beam", beam_width=1)
    print(scale_name, "greedy bf16==fp32:", sum(_prefix([1] + list(x[1:])) == _prefix([1] + list(y[1:])) for x, y in zip(g32, gbf)),
          "beam5 bf16==fp32:", sum(_prefix(x) == _prefix(y) for x, y in zip(a, b)),
          "beam1==greedy (bf16):", sum(_prefix([1] + list(x[1:]))[1:] == _prefix(y)[1:len(_prefix([1] + list(x[1:])))] for x, y in zip(gbf, b1)), "of", B)
