"""tda_eeg_audio_b200 — B200-native engine for the windowed TDA feature path of
Ignaciagothe/tda-eeg-audio.  Hand-written sm_100a CUDA behind a C-ABI
(include/tda_b200.h); this package is the thin Python mirror of the reference's call surface."""
from . import _lib  # noqa: F401
from .rips import rips_h01_batched, ripser  # noqa: F401

__all__ = ["rips_h01_batched", "ripser"]
