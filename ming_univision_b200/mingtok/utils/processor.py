"""`mingtok.utils.CenterCropProcessor` (mingtok/utils/processor.py:7-46) on the GPU.

The reference composes torchvision CPU transforms on a PIL image — Resize(image_size, BICUBIC) -> CenterCrop(image_size)
-> ToTensor -> Normalize(mean, std) — and the caller then moves the fp32 tensor to the device
(mingunivision/test_infer_recon_image.py:19-20).  Here the DECODED u8 pixels go to the device (3 bytes per input pixel
from pinned memory instead of 12 bytes per output value) and one C-ABI call (`mb_image_preprocess_u8`) does the whole
transform there, bit-exactly: Pillow's antialiased bicubic resample in its 8-bit fixed point (horizontal pass, u8
rounding, vertical pass), torchvision's crop offsets, `/ 255`, `(x - mean) / std`.  Same constructor, `__call__` and
`from_config`; the result is the same [3, S, S] fp32 tensor, already on the device (`.cuda()` on it is a no-op).

`ResizeProcessor` is the no-crop variant of mingunivision/processing_bailingmm.py:80-100 (`MingTokUndProcessor`:
Resize((S, S)) without keeping the aspect ratio); `MingTokCenterCropProcessor` there (:102-123) is `CenterCropProcessor`
with CLIP's default statistics.  `tensor_to_pil` is modeling_bailing_moe.py:84-90.

No CPU fallback: the transforms need the CUDA library (they raise without a device).
"""
from __future__ import annotations

import numpy as np
import torch

from ... import ops

_HALF = (0.5, 0.5, 0.5)
_CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
_CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


def _to_u8_hwc(item, device) -> torch.Tensor:
    """PIL image / numpy [H, W, 3] u8 / torch uint8 [H, W, 3] -> CUDA uint8 [H, W, 3] (pinned staging for host data)."""
    if isinstance(item, torch.Tensor):
        t = item
    else:
        if hasattr(item, "convert"):  # PIL.Image
            if item.mode != "RGB":
                raise ValueError(f"expected an RGB image, got mode {item.mode} (the reference converts first, "
                                 "test_infer_recon_image.py:17)")
            arr = np.array(item)  # a writable copy of the decoded pixels
        else:
            arr = np.asarray(item)
        if arr.dtype != np.uint8 or arr.ndim != 3 or arr.shape[2] != 3:
            raise ValueError(f"expected [H, W, 3] uint8 pixels, got {arr.dtype} {arr.shape}")
        t = torch.from_numpy(np.ascontiguousarray(arr))
    if t.dtype != torch.uint8 or t.dim() != 3 or t.shape[2] != 3:
        raise ValueError(f"expected [H, W, 3] uint8 pixels, got {t.dtype} {tuple(t.shape)}")
    if not t.is_cuda:
        t = t.pin_memory().to(device, non_blocking=True)
    return t


class CenterCropProcessor:
    def __init__(self, image_size=512, mean=None, std=None, device="cuda", dtype=torch.float32):
        self.image_size = int(image_size)
        self.mean = tuple(float(v) for v in (_HALF if mean is None else mean))
        self.std = tuple(float(v) for v in (_HALF if std is None else std))
        self.device = torch.device(device)
        self.dtype = dtype

    def __call__(self, item) -> torch.Tensor:
        src = _to_u8_hwc(item, self.device)
        return ops.image_preprocess(src, self.image_size, self.image_size, self.mean, self.std, self.dtype)[0]

    def batch(self, images: torch.Tensor) -> torch.Tensor:
        """Same-size images at once: uint8 [N, H, W, 3] on the device -> [N, 3, S, S]."""
        return ops.image_preprocess(images, self.image_size, self.image_size, self.mean, self.std, self.dtype)

    @classmethod
    def from_config(cls, cfg=None):
        cfg = {} if cfg is None else cfg
        return cls(image_size=cfg.get("image_size", 512), mean=cfg.get("mean", None), std=cfg.get("std", None))


class ResizeProcessor:
    """Resize((S, S), BICUBIC) -> ToTensor -> Normalize: processing_bailingmm.py:80-100 (CLIP statistics by default, as
    there; the reference instantiates it with 0.5 / 0.5 at :175)."""

    def __init__(self, image_size=224, mean=None, std=None, device="cuda", dtype=torch.float32):
        self.image_size = int(image_size)
        self.mean = tuple(float(v) for v in (_CLIP_MEAN if mean is None else mean))
        self.std = tuple(float(v) for v in (_CLIP_STD if std is None else std))
        self.device = torch.device(device)
        self.dtype = dtype

    def __call__(self, item) -> torch.Tensor:
        src = _to_u8_hwc(item, self.device)
        return ops.image_preprocess(src, (self.image_size, self.image_size), None, self.mean, self.std, self.dtype)[0]


def tensor_to_pil(image_tensor: torch.Tensor, mean=_HALF, std=_HALF):
    """modeling_bailing_moe.py:84-90: [1, 3, H, W] (or [3, H, W]) in [-1, 1] -> PIL image; the u8 conversion (x*std + mean,
    * 255, truncation) runs on the device and only H*W*3 bytes cross PCIe."""
    from PIL import Image

    u8 = ops.image_postprocess(image_tensor if image_tensor.dim() == 3 else image_tensor[0], mean, std)[0]
    return Image.fromarray(u8.cpu().numpy())
