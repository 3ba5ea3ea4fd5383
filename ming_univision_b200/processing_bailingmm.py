"""The two image transforms of `processing_bailingmm.py` under their reference names, on the GPU:

    MingTokUndProcessor(image_size=224, mean=None, std=None)         processing_bailingmm.py:80-100
        Resize((S, S), BICUBIC) -> ToTensor -> Normalize             (understanding: S = 1024, :175)
    MingTokCenterCropProcessor(image_size=224, mean=None, std=None)  processing_bailingmm.py:102-123
        Resize(S, BICUBIC) -> CenterCrop(S) -> ToTensor -> Normalize (generation / editing: S = 512, :176)

Both default to CLIP's statistics, as there; `BailingMMProcessor.__init__` passes 0.5 / 0.5.  They are the classes of
`ming_univision_b200.mingtok.utils.processor` (one C-ABI call, `mb_image_preprocess_u8`, bit-exact with the PIL /
torchvision stack); the rest of `BailingMMProcessor` (chat template, token expansion, tokenizer) is host-side string
processing and stays the reference's own code.  INTEGRATION.md shows the three lines a maintainer
changes in `processing_bailingmm.py` to use them.
"""
from __future__ import annotations

from .mingtok.utils.processor import _CLIP_MEAN, _CLIP_STD, CenterCropProcessor, ResizeProcessor


class MingTokUndProcessor(ResizeProcessor):
    pass


class MingTokCenterCropProcessor(CenterCropProcessor):
    def __init__(self, image_size=224, mean=None, std=None, **kw):
        super().__init__(image_size=image_size, mean=_CLIP_MEAN if mean is None else mean,
                         std=_CLIP_STD if std is None else std, **kw)


def install_gpu_image_processors(processor, image_size_und: int = 1024, image_size_gen: int = 512):
    """Swaps the two torchvision stacks of a live reference `BailingMMProcessor` (attributes `vis_processor` /
    `gen_processor`, processing_bailingmm.py:175-176) for the device versions.  The reference's `__call__` still has to
    gather the per-image tensors with `torch.stack` instead of `np.array` (:264) — see INTEGRATION.md."""
    half = [0.5, 0.5, 0.5]
    processor.vis_processor = MingTokUndProcessor(image_size=image_size_und, mean=half, std=half)
    processor.gen_processor = MingTokCenterCropProcessor(image_size=image_size_gen, mean=half, std=half)
    return processor
