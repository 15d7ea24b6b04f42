"""turtlevsr_b200 -- B200-native implementation of Turtle's inference hot path.

Host side (Python) mirrors the reference arch interface; the computation is hand-written
sm_100a CUDA behind a C-ABI (include/turtle_b200.h, turtlevsr_b200/csrc).
"""
from .archs import create_video_model  # noqa: F401

__all__ = ["create_video_model"]
__version__ = "0.1.0"
