"""ming_univision_b200 — B200-native (sm_100a) drop-in for the continuous-visual-token hot path of
inclusionAI/Ming-UniVision: MingTok encoder / semantic decoder / pixel decoder, the Bailing-MoE AR step and the
rectified-flow SwiGLU head.  Host code is Python/PyTorch (device memory, streams, torch.distributed); every operator
is a hand-written CUDA kernel reached through the C ABI declared in ``include/mingb200.h`` (``libmingb200.so``).

There is no CPU or eager-PyTorch fallback: importing the package works anywhere (so CPU-only tooling can inspect
it), but calling an operator without the built library or without an sm_100 device raises ``RuntimeError``.
"""
from . import _lib  # noqa: F401

__version__ = "0.1.0"

# The reference's scripts import its modules FLAT after putting `mingunivision/` on sys.path
# (mingunivisioninfer.py:1-7: `from modeling_bailingmm import MingUniVisionForConditionalGeneration`; app.py,
# test_infer_unified.py: `from mingunivisioninfer import MingUniVisionInfer`; modeling_bailingmm.py:25:
# `from mingtok.modeling_mingtok import MingTok`).  One call makes those unchanged import lines resolve to this package.
_FLAT_MODULES = ("modeling_bailingmm", "modeling_bailing_moe", "diff_loss_rf_swiglu", "mingunivisioninfer",
                 "image_processing_bailingmm", "processing_bailingmm", "mingtok", "mingtok.modeling_mingtok",
                 "mingtok.utils", "mingtok.utils.processor")


def install_flat_modules(force: bool = False) -> list:
    """Registers this package's modules under the reference's flat module names in `sys.modules` (`modeling_bailingmm`,
    `modeling_bailing_moe`, `diff_loss_rf_swiglu`, `mingunivisioninfer`, `image_processing_bailingmm`,
    `processing_bailingmm`, `mingtok.*`), so a
    reference script runs on the B200-native path with ONE added line at its top.  Names that are already imported (a
    reference checkout earlier on sys.path) are left alone unless `force`.  Returns the names it registered."""
    import importlib
    import sys

    done = []
    for name in _FLAT_MODULES:
        if name in sys.modules and not force:
            continue
        sys.modules[name] = importlib.import_module(f"{__name__}.{name}")
        done.append(name)
    return done
