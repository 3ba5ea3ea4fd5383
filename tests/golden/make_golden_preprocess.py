"""Golden vectors of the image pre- / post-processing, produced by the REAL third-party libraries the
reference calls — Pillow's `Image.resize(BICUBIC)` and torchvision's Resize / CenterCrop / ToTensor / Normalize /
ToPILImage, composed exactly as mingtok/utils/processor.py:17-27 and mingunivision/processing_bailingmm.py:80-123 do —
run in the build container (Pillow 12.2.0, torchvision 0.26.0):

    python tests/golden/make_golden_preprocess.py        -> tests/golden/preprocess.npz

Inputs are seeded synthetic photographs (low-pass noise + grain), stored with their outputs so that the fixtures do not
depend on any generator's stream.
"""
import json
import os

import numpy as np
import torch
import torchvision.transforms as T
from PIL import Image
from torchvision.transforms import InterpolationMode

HERE = os.path.dirname(os.path.abspath(__file__))

# (name, H, W, size, crop, mean, std)
HALF = (0.5, 0.5, 0.5)
CLIP_MEAN, CLIP_STD = (0.48145466, 0.4578275, 0.40821073), (0.26862954, 0.26130258, 0.27577711)
CASES = [
    ("portrait_down", 131, 97, 64, 64, HALF, HALF),          # CenterCropProcessor: rows cropped
    ("landscape_down", 90, 160, 48, 48, HALF, HALF),         # columns cropped
    ("upscale", 40, 56, 96, 96, HALF, HALF),                 # enlarging: support stays 2
    ("und_square", 75, 120, (64, 64), None, HALF, HALF),     # MingTokUndProcessor: aspect ratio not kept
    ("clip_stats", 100, 100, 32, 32, CLIP_MEAN, CLIP_STD),   # the classes' default statistics
    ("same_size", 64, 64, 64, 64, HALF, HALF),               # Resize returns the image untouched
    ("width_only", 64, 100, (64, 50), None, HALF, HALF),     # Pillow skips the vertical pass
    ("height_only", 100, 64, (50, 64), None, HALF, HALF),    # ... and the horizontal one
]


def synthetic_photo(rng, h, w):
    base = rng.integers(0, 256, (h // 8 + 2, w // 8 + 2, 3), dtype=np.uint8)
    img = np.asarray(Image.fromarray(base).resize((w, h), Image.BILINEAR)).astype(np.int32)
    return np.clip(img + rng.integers(-24, 25, (h, w, 3)), 0, 255).astype(np.uint8)


def main():
    rng = np.random.default_rng(20240917)
    out = {}
    for name, h, w, size, crop, mean, std in CASES:
        img = synthetic_photo(rng, h, w)
        tf = [T.Resize(size=size, interpolation=InterpolationMode.BICUBIC)]
        if crop is not None:
            tf.append(T.CenterCrop(crop))
        u8 = np.asarray(T.Compose(tf)(Image.fromarray(img)))
        ten = T.Compose(tf + [T.ToTensor(), T.Normalize(mean, std)])(Image.fromarray(img)).numpy()
        out[f"{name}.src"] = img
        out[f"{name}.u8"] = u8
        out[f"{name}.tensor"] = ten
    # post-processing: tensor_to_pil (modeling_bailing_moe.py:84-90) on values around the 1/255 steps
    x = torch.from_numpy(rng.uniform(-1, 1, (1, 3, 24, 40)).astype(np.float32))
    x[0, 0, 0, :6] = torch.tensor([-1.0, 1.0, 0.0, 2 / 255 - 1, 0.99999, -0.99999])
    half = torch.tensor(HALF).view(1, -1, 1, 1)
    out["post.x"] = x.numpy()
    out["post.u8"] = np.asarray(T.ToPILImage()((x * half + half)[0]))
    out["cases"] = np.array(json.dumps([dict(name=c[0], size=c[3], crop=c[4], mean=c[5], std=c[6]) for c in CASES]))
    np.savez_compressed(os.path.join(HERE, "preprocess.npz"), **out)
    print("wrote", os.path.join(HERE, "preprocess.npz"), os.path.getsize(os.path.join(HERE, "preprocess.npz")), "bytes")


if __name__ == "__main__":
    main()
