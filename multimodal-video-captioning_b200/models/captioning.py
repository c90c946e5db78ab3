"""models.captioning (reference src/models/captioning.py): same public names, B200 arithmetic."""
from salstm.modules import (AVCaptioning, AVCaptioningDual, DECODER_CONFIG, RECONSTRUCTOR_CONFIG,  # noqa: F401
                            VISUAL_DECODER_CONFIG, AUDIO_DECODER_CONFIG)
