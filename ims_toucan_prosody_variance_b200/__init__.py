"""toucan_b200 -- B200-native engine for the IMS-Toucan text->wave hot path.

Host code is Python/PyTorch (device memory, streams); every hot op is a hand-written sm_100a
kernel in libtoucan_b200.so behind the C ABI of include/toucan_b200.h.  No CPU fallback.
"""
from . import _lib, layouts, ops, sharding  # noqa: F401
from .frontend import PhoneTensoriser  # noqa: F401
from .interface import ToucanTTSInterface, UtteranceCloner  # noqa: F401
from .pipeline import TextToWave  # noqa: F401
from .toucantts import ToucanTTS  # noqa: F401
from .vocoder import BigVGAN, HiFiGANGenerator  # noqa: F401

__all__ = ["ToucanTTS", "BigVGAN", "HiFiGANGenerator", "TextToWave", "ToucanTTSInterface", "UtteranceCloner", "PhoneTensoriser", "ops", "layouts",
           "sharding"]
