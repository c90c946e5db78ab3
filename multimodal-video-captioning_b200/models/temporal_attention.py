"""models.temporal_attention (reference src/models/temporal_attention.py)."""
from salstm.modules import TemporalAttention  # noqa: F401
