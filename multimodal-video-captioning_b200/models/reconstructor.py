"""models.reconstructor (reference src/models/reconstructor.py)."""
from salstm.modules import GlobalReconstructor, LocalReconstructor, build_caption_mask  # noqa: F401
