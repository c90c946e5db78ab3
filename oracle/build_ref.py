"""Recipe for oracle/_ref: the UNMODIFIED hot-path sources of the reference, staged for the CPU baseline.

TEST / MEASUREMENT INFRASTRUCTURE ONLY (like everything under oracle/): imported by tests/, by
__graft_entry__.build()/smoke() and by bench.py's `--impl reference` / cpu_baseline legs, never by the
product package.

The reference is a flat script tree (no setup.py / pyproject), so "building" it is staging: when
/root/reference is present (the build container), `build()` copies src/models/*.py and src/losses.py
byte for byte into oracle/_ref/src/ and records their sha256 in oracle/_ref/MANIFEST.json.  oracle/_ref/
is git-ignored (reference sources never enter the history) but NOT gpurun-ignored, so the staged files
travel to the GPU box, where /root/reference does not exist.

`load()` imports the staged modules WITHOUT executing models/__init__.py (src/models/__init__.py:4-5 pulls in
the offline feature encoders: torch.hub / torchvision) by registering a bare namespace package `models`
whose __path__ is the staged directory, and with placeholder `pycocoevalcap.*` modules, which src/losses.py
imports at module level (:6-9) only for NLPScore (string metrics; Java for METEOR).  Nothing in the staged
files is edited.
"""
from __future__ import annotations

import hashlib
import importlib
import importlib.util
import json
import os
import shutil
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("MVC_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(HERE, "_ref")
FILES = ["src/models/__init__.py", "src/models/temporal_attention.py", "src/models/features_captioning.py",
         "src/models/reconstructor.py", "src/models/captioning.py", "src/losses.py",
         # the two callers either side of the path, for the train.py-level drop-in test on the GPU box
         # (tests/test_gpu_dropin.py runs the reference's Trainer over the B200 modules); never imported by load()
         "src/train.py", "src/get_loader.py"]


def _sha(path):
    with open(path, "rb") as fh:
        return hashlib.sha256(fh.read()).hexdigest()


def build() -> bool:
    """Stage the reference sources (no-op without /root/reference).  Returns True when oracle/_ref is usable."""
    if os.path.isdir(os.path.join(REF_ROOT, "src", "models")):
        manifest = {}
        for rel in FILES:
            src, dst = os.path.join(REF_ROOT, rel), os.path.join(OUT, rel)
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            shutil.copyfile(src, dst)
            manifest[rel] = _sha(dst)
        with open(os.path.join(OUT, "MANIFEST.json"), "w") as fh:
            json.dump({"source": "hmartelb/multimodal-video-captioning (unmodified copies)", "sha256": manifest}, fh,
                      indent=1)
    return available()


def available() -> bool:
    man = os.path.join(OUT, "MANIFEST.json")
    if not os.path.isfile(man):
        return False
    try:
        sha = json.load(open(man))["sha256"]
        return all(os.path.isfile(os.path.join(OUT, rel)) and _sha(os.path.join(OUT, rel)) == h for rel, h in sha.items())
    except Exception:
        return False


class _Saved:
    """Temporarily own sys.modules['models'/'losses'/'pycocoevalcap*'] so the product's same-named modules
    (multimodal-video-captioning_b200/models, losses.py) and the staged reference never mix."""
    NAMES = ("models", "losses", "pycocoevalcap")

    def __enter__(self):
        self.saved = {k: v for k, v in sys.modules.items() if k.split(".")[0] in self.NAMES}
        for k in self.saved:
            del sys.modules[k]
        return self

    def __exit__(self, *exc):
        for k in [k for k in sys.modules if k.split(".")[0] in self.NAMES]:
            del sys.modules[k]
        sys.modules.update(self.saved)
        return False


_CACHE = None


def load():
    """-> namespace with the reference's AVCaptioning, AVCaptioningDual, FeaturesCaptioning, GlobalReconstructor,
    LocalReconstructor, TemporalAttention classes and its `losses` module, executed from oracle/_ref."""
    global _CACHE
    if _CACHE is not None:
        return _CACHE
    if not available():
        raise RuntimeError("oracle/_ref is not staged (run oracle/build_ref.py where /root/reference exists)")
    src = os.path.join(OUT, "src")
    with _Saved():
        pkg = types.ModuleType("models")
        pkg.__path__ = [os.path.join(src, "models")]          # namespace-style: submodules importable, __init__ not run
        sys.modules["models"] = pkg
        for name, attrs in (("pycocoevalcap", ()), ("pycocoevalcap.bleu", ()), ("pycocoevalcap.bleu.bleu", ("Bleu",)),
                            ("pycocoevalcap.rouge", ()), ("pycocoevalcap.rouge.rouge", ("Rouge",)),
                            ("pycocoevalcap.cider", ()), ("pycocoevalcap.cider.cider", ("Cider",)),
                            ("pycocoevalcap.meteor", ()), ("pycocoevalcap.meteor.meteor", ("Meteor",))):
            m = types.ModuleType(name)
            m.__path__ = []
            for a in attrs:
                setattr(m, a, None)                            # only NLPScore touches these; it is never called here
            sys.modules[name] = m
        cap = importlib.import_module("models.captioning")
        fc = importlib.import_module("models.features_captioning")
        rec = importlib.import_module("models.reconstructor")
        ta = importlib.import_module("models.temporal_attention")
        spec = importlib.util.spec_from_file_location("losses", os.path.join(src, "losses.py"))
        losses = importlib.util.module_from_spec(spec)
        sys.modules["losses"] = losses
        spec.loader.exec_module(losses)
    _CACHE = types.SimpleNamespace(AVCaptioning=cap.AVCaptioning, AVCaptioningDual=cap.AVCaptioningDual,
                                   FeaturesCaptioning=fc.FeaturesCaptioning, GlobalReconstructor=rec.GlobalReconstructor,
                                   LocalReconstructor=rec.LocalReconstructor, TemporalAttention=ta.TemporalAttention,
                                   losses=losses, root=src)
    return _CACHE


if __name__ == "__main__":
    print("oracle/_ref staged:", build())
