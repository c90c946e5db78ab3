"""Drop-in for the reference's `models` package (src/models/__init__.py:1-5): put
`multimodal-video-captioning_b200/` ahead of the reference's `src/` on sys.path and
`from models import AVCaptioning, AVCaptioningDual` (train.py:13) resolves here.
AudioEncoder / VisualEncoder are the offline feature precompute and stay the reference's."""
from .features_captioning import FeaturesCaptioning
from .captioning import AVCaptioning
from .captioning import AVCaptioningDual
from .temporal_attention import TemporalAttention
from .reconstructor import GlobalReconstructor, LocalReconstructor, build_caption_mask

__all__ = ["FeaturesCaptioning", "AVCaptioning", "AVCaptioningDual", "TemporalAttention", "GlobalReconstructor",
           "LocalReconstructor", "build_caption_mask"]
