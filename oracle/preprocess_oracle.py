"""CPU oracle for the image pre- / post-processing either side of the MingTok path:

    pre :  PIL image -> Resize(bicubic) [-> CenterCrop] -> ToTensor -> Normalize      mingtok/utils/processor.py:17-27,
                                                            mingunivision/processing_bailingmm.py:80-123 (both variants)
    post:  [-1, 1] CHW tensor -> x*std + mean -> ToPILImage (mul(255).byte(): TRUNCATION)
                                                            mingunivision/modeling_bailing_moe.py:84-90, :1787
                                                            mingunivision/test_infer_recon_image.py:24-28

TEST INFRASTRUCTURE — NOT PRODUCT CODE (see oracle/mingtok_oracle.py for the rules).

The arithmetic lives in third-party code that is NOT under /root/reference:
  * Pillow (unpinned in the reference's requirements.txt; 12.2.0 in this image): `Image.resize(BICUBIC)` =
    src/libImaging/Resample.c — `precompute_coeffs` (antialiased: filter support scaled by the down-scale factor, Keys
    cubic a = -0.5, weights normalised in double), `normalize_coeffs_8bpc` (fixed point, 22 fractional bits, round half
    away from zero), `ImagingResampleHorizontal_8bpc` then `ImagingResampleVertical_8bpc` (int32 accumulation from
    1 << 21, arithmetic shift, clip to u8 BETWEEN the two passes); PIL/Image.py `Image.resize` (the pass order flips for
    images more than 100 times taller than wide).  Restated below in numpy from its published algorithm.
  * torchvision 0.26 (`torchvision==0.22.0` in requirements.txt:2): `_compute_resized_output_size` (short edge -> size,
    long edge = int(size * long / short)), `center_crop` (offset = int(round((full - crop) / 2.0)), Python's half-to-even),
    `to_tensor` (u8 -> fp32, true division by 255), `normalize` ((x - mean) / std in fp32).
Pinned: tests/test_preprocess_cpu.py compares every function here with Pillow + torchvision THEMSELVES, run in this
container, bit-exactly (u8 images) / exactly (fp32 tensors) on seeded images over up- and down-scales, odd sizes and the
reference's three processor configurations; tests/golden/preprocess.npz holds outputs of the real libraries for the GPU
box (made by tests/golden/make_golden_preprocess.py).
"""
from __future__ import annotations

import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2  # Resample.c: 8 bits for the result, 2 of headroom for the cubic's over/undershoot
BICUBIC_SUPPORT = 2.0


def bicubic_filter(x: float) -> float:
    """Resample.c `bicubic_filter` (Keys, a = -0.5), the exact operation order in double."""
    a = -0.5
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def precompute_coeffs(in_size: int, out_size: int) -> tuple[np.ndarray, np.ndarray, int]:
    """Resample.c `precompute_coeffs` + `normalize_coeffs_8bpc` for the full box [0, in_size).
    Returns (bounds int32 [out, 2] = (xmin, count), kk int32 [out, ksize], ksize)."""
    scale = float(np.float32(in_size) - np.float32(0.0)) / out_size
    filterscale = scale if scale >= 1.0 else 1.0
    support = BICUBIC_SUPPORT * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)  # C cast: truncation toward zero
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = [bicubic_filter((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        for x in range(xmax):
            v = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk, ksize


def _resample_axis1(img: np.ndarray, bounds: np.ndarray, kk: np.ndarray) -> np.ndarray:
    """One 8-bit pass along axis 1 of img [A, in, C] u8 -> [A, out, C] u8 (ImagingResampleHorizontal_8bpc; the vertical
    pass is the same arithmetic along the other axis)."""
    out_size, ksize = kk.shape
    idx = bounds[:, :1].astype(np.int64) + np.arange(ksize)[None]  # [out, ksize]; taps past `count` have weight 0
    idx = np.minimum(idx, img.shape[1] - 1)
    out = np.empty((img.shape[0], out_size, img.shape[2]), dtype=np.uint8)
    step = max(1, (1 << 24) // max(1, img.shape[0] * ksize * img.shape[2]))
    for o0 in range(0, out_size, step):
        o1 = min(out_size, o0 + step)
        taps = img[:, idx[o0:o1]].astype(np.int64)  # [A, o, ksize, C]
        acc = (taps * kk[o0:o1].astype(np.int64)[None, :, :, None]).sum(axis=2) + (1 << (PRECISION_BITS - 1))
        assert np.abs(acc).max() < 2 ** 31, "the C code accumulates in int32"
        out[:, o0:o1] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return out


def resize_bicubic_u8(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """PIL `Image.resize((out_w, out_h), BICUBIC)` on an RGB / L image: img [H, W, C] u8 -> [out_h, out_w, C] u8.
    ImagingResample: horizontal pass first (only if the width changes), then vertical (only if the height changes)."""
    if img.dtype != np.uint8 or img.ndim != 3:
        raise ValueError("expected a [H, W, C] uint8 image")
    h, w, _ = img.shape

    def horizontal(a):
        if a.shape[1] == out_w:
            return a
        bh, kh, _ = precompute_coeffs(a.shape[1], out_w)
        return _resample_axis1(a, bh, kh)

    def vertical(a):
        if a.shape[0] == out_h:
            return a
        bv, kv, _ = precompute_coeffs(a.shape[0], out_h)
        return _resample_axis1(a.transpose(1, 0, 2), bv, kv).transpose(1, 0, 2)

    if vertical_pass_first(h, w, out_h):
        img = horizontal(np.ascontiguousarray(vertical(img)))
    else:
        img = vertical(horizontal(img))
    return np.ascontiguousarray(img)


def vertical_pass_first(h: int, w: int, out_h: int) -> bool:
    """Pillow >= 11 (`Image.resize`, PIL/Image.py: `if self.size[1] > self.size[0] * 100 and size[1] < self.size[1]`)
    resizes images more than 100 times taller than wide in two calls, height first, when the height shrinks; the u8
    rounding between the passes makes the order visible in the result.  Found by fuzzing the oracle against Pillow 12.2;
    photographs never get there, and the device path refuses the geometry (MB_ERR_SHAPE) instead of guessing."""
    return h > w * 100 and out_h < h


def resized_output_size(h: int, w: int, size) -> tuple[int, int]:
    """torchvision `_compute_resized_output_size` without max_size: int -> short edge, (h, w) -> exact."""
    if isinstance(size, (tuple, list)):
        if len(size) == 2:
            return int(size[0]), int(size[1])
        size = size[0]
    short, long = (w, h) if w <= h else (h, w)
    new_short, new_long = size, int(size * long / short)
    new_w, new_h = (new_short, new_long) if w <= h else (new_long, new_short)
    return new_h, new_w


def center_crop_offsets(h: int, w: int, crop_h: int, crop_w: int) -> tuple[int, int]:
    """torchvision `center_crop` for crop <= image (Python's round: half to even)."""
    if crop_h > h or crop_w > w:
        raise ValueError("crop larger than the image (torchvision would zero-pad; the path never does this)")
    return int(round((h - crop_h) / 2.0)), int(round((w - crop_w) / 2.0))


def to_tensor_normalize(img: np.ndarray, mean, std) -> np.ndarray:
    """`ToTensor` + `Normalize`: [H, W, 3] u8 -> [3, H, W] fp32 = ((x / 255) - mean) / std, every step rounded to fp32."""
    x = img.astype(np.float32) / np.float32(255.0)
    m = np.asarray(mean, dtype=np.float32)
    s = np.asarray(std, dtype=np.float32)
    return np.ascontiguousarray(((x - m) / s).astype(np.float32).transpose(2, 0, 1))


def preprocess(img: np.ndarray, size, crop: int | None, mean=(0.5, 0.5, 0.5), std=(0.5, 0.5, 0.5)) -> np.ndarray:
    """The whole transform of `CenterCropProcessor` (size=int, crop=size), `MingTokCenterCropProcessor` (same) and
    `MingTokUndProcessor` (size=(s, s), crop=None): [H, W, 3] u8 -> [3, h, w] fp32."""
    rh, rw = resized_output_size(img.shape[0], img.shape[1], size)
    out = resize_bicubic_u8(img, rh, rw)
    if crop is not None:
        top, left = center_crop_offsets(rh, rw, crop, crop)
        out = out[top:top + crop, left:left + crop]
    return to_tensor_normalize(out, mean, std)


def postprocess(x: np.ndarray, mean=(0.5, 0.5, 0.5), std=(0.5, 0.5, 0.5)) -> np.ndarray:
    """`tensor_to_pil` (modeling_bailing_moe.py:84-90): [3, H, W] fp32 -> [H, W, 3] u8 = trunc((x*std + mean) * 255);
    fp32 steps, and `Tensor.byte()` truncates toward zero (values outside [0, 255] wrap there; the path clamps to
    [-1, 1] first, modeling_mingtok.py:194, so they do not occur and the oracle saturates instead)."""
    x = x.astype(np.float32)
    m = np.asarray(mean, dtype=np.float32)[:, None, None]
    s = np.asarray(std, dtype=np.float32)[:, None, None]
    y = ((x * s).astype(np.float32) + m).astype(np.float32) * np.float32(255.0)
    return np.ascontiguousarray(np.clip(np.trunc(y), 0, 255).astype(np.uint8).transpose(1, 2, 0))
