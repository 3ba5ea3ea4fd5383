"""ming_univision_b200 — B200-native (sm_100a) drop-in for the continuous-visual-token hot path of
inclusionAI/Ming-UniVision: MingTok encoder / semantic decoder / pixel decoder, the Bailing-MoE AR step and the
rectified-flow SwiGLU head.  Host code is Python/PyTorch (device memory, streams, torch.distributed); every operator
is a hand-written CUDA kernel reached through the C ABI declared in ``include/mingb200.h`` (``libmingb200.so``).

There is no CPU or eager-PyTorch fallback: importing the package works anywhere (so CPU-only tooling can inspect
it), but calling an operator without the built library or without an sm_100 device raises ``RuntimeError``.
"""
from . import _lib  # noqa: F401

__version__ = "0.1.0"
