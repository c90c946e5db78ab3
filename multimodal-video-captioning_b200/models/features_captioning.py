"""models.features_captioning (reference src/models/features_captioning.py)."""
from salstm.modules import FeaturesCaptioning  # noqa: F401
