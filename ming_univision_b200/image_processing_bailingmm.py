"""`BailingMMImageProcessor` — the surface of the reference's ``mingunivision/image_processing_bailingmm.py`` (:94-122
`smart_resize`, :124-462 the processor class): dynamic-resolution resize to multiples of patch x merge inside a pixel budget,
rescale, normalise, and the flattened (temporal, merge-window, patch) layout with its `(t, h, w)` grid.

In Ming-UniVision this processor serves the VIDEO inputs of `BailingMMProcessor` (processing_bailingmm.py:272); still images
go through the MingTok processors (`mingtok.utils.processor`, device kernels in csrc/preprocess.cu).  It is host-side
pre-processing ("unchanged CPU code"), so this is a host implementation: PIL does the resample exactly as the
reference's `transformers.image_transforms.resize` does, numpy the arithmetic and the patch re-layout.  `device=` moves the
bicubic resize of uint8 frames onto the GPU (`ops.image_resize_u8`, bit-identical to Pillow — tests/test_preprocess_*).
"""
from __future__ import annotations

import math
from typing import List, Optional, Union

import numpy as np
from transformers.feature_extraction_utils import BatchFeature
from transformers.image_processing_utils import BaseImageProcessor
from transformers.image_transforms import convert_to_rgb, resize, to_channel_dimension_format
from transformers.image_utils import (OPENAI_CLIP_MEAN, OPENAI_CLIP_STD, ChannelDimension, PILImageResampling,
                                      get_image_size, infer_channel_dimension_format, is_valid_image,
                                      make_list_of_images, to_numpy_array)


def smart_resize(height: int, width: int, factor: int = 28, min_pixels: int = 56 * 56,
                 max_pixels: int = 14 * 14 * 4 * 1280) -> tuple[int, int]:
    """Target size (image_processing_bailingmm.py:94-122): both sides multiples of `factor`, the area inside
    [min_pixels, max_pixels], the aspect ratio kept as well as that allows."""
    if min(height, width) < factor:
        raise ValueError(f"height:{height} or width:{width} must be larger than factor:{factor}")
    ratio = max(height, width) / min(height, width)
    if ratio > 200:
        raise ValueError(f"absolute aspect ratio must be smaller than 200, got {ratio}")
    h, w = round(height / factor) * factor, round(width / factor) * factor
    if h * w > max_pixels:      # shrink: floor, so the budget is never exceeded
        beta = math.sqrt(height * width / max_pixels)
        h, w = math.floor(height / beta / factor) * factor, math.floor(width / beta / factor) * factor
    elif h * w < min_pixels:    # grow: ceil, so the minimum is always reached
        beta = math.sqrt(min_pixels / (height * width))
        h, w = math.ceil(height * beta / factor) * factor, math.ceil(width * beta / factor) * factor
    return h, w


def _as_image_batches(images) -> list:
    """A single image, a list of images, or a list of lists -> list of per-sample frame lists (one frame each)."""
    if isinstance(images, (list, tuple)) and images and isinstance(images[0], (list, tuple)):
        return [img for sample in images for img in sample]
    if isinstance(images, (list, tuple)):
        return list(images)
    if is_valid_image(images):
        return [images]
    raise ValueError(f"Could not make batched images from {type(images)}")


def _as_video_batches(videos) -> list:
    """-> list of videos, each a list of frames (a 4-D array is one video; a flat list of frames is one video)."""
    if isinstance(videos, (list, tuple)) and videos and isinstance(videos[0], (list, tuple)):
        return [list(v) for v in videos]
    if isinstance(videos, (list, tuple)) and videos and is_valid_image(videos[0]):
        first = videos[0]
        if hasattr(first, "ndim") and first.ndim == 4:
            return [list(v) for v in videos]
        return [list(videos)]
    if hasattr(videos, "ndim") and videos.ndim == 4:
        return [list(videos)]
    raise ValueError(f"Could not make batched video from {type(videos)}")


class BailingMMImageProcessor(BaseImageProcessor):
    model_input_names = ["pixel_values", "image_grid_thw", "pixel_values_videos", "video_grid_thw"]

    def __init__(self, do_resize: bool = True, resample: PILImageResampling = PILImageResampling.BICUBIC,
                 do_rescale: bool = True, rescale_factor: Union[int, float] = 1 / 255, do_normalize: bool = True,
                 image_mean: Optional[Union[float, List[float]]] = None,
                 image_std: Optional[Union[float, List[float]]] = None, do_convert_rgb: bool = True,
                 min_pixels: int = 56 * 56, max_pixels: int = 28 * 28 * 1280, min_pixels_video: int = 128 * 28 * 28,
                 max_pixels_video: int = 768 * 28 * 28, patch_size: int = 14, temporal_patch_size: int = 2,
                 merge_size: int = 2, **kwargs) -> None:
        super().__init__(**kwargs)
        self.do_resize, self.resample = do_resize, resample
        self.do_rescale, self.rescale_factor = do_rescale, rescale_factor
        self.do_normalize = do_normalize
        self.image_mean = image_mean if image_mean is not None else OPENAI_CLIP_MEAN
        self.image_std = image_std if image_std is not None else OPENAI_CLIP_STD
        self.min_pixels, self.max_pixels = min_pixels, max_pixels
        self.min_pixels_video, self.max_pixels_video = min_pixels_video, max_pixels_video
        self.patch_size, self.temporal_patch_size, self.merge_size = patch_size, temporal_patch_size, merge_size
        self.size = {"min_pixels": min_pixels, "max_pixels": max_pixels}
        self.do_convert_rgb = do_convert_rgb

    # -- one sample (an image = one frame, or the frames of one video) -> flattened patches + (t, h, w) grid ----------
    def _preprocess(self, images, do_resize=None, resample=None, do_rescale=None, rescale_factor=None, do_normalize=None,
                    image_mean=None, image_std=None, do_convert_rgb=None, data_format=ChannelDimension.FIRST,
                    input_data_format=None, min_pixels=None, max_pixels=None, device=None):
        frames = make_list_of_images(images)
        if do_convert_rgb:
            frames = [convert_to_rgb(f) for f in frames]
        frames = [to_numpy_array(f) for f in frames]
        if input_data_format is None:
            input_data_format = infer_channel_dimension_format(frames[0])
        height, width = get_image_size(frames[0], channel_dim=input_data_format)
        out_h, out_w = height, width
        if do_resize:  # every frame of the sample gets the size computed from the FIRST frame (as the reference does)
            out_h, out_w = smart_resize(height, width, factor=self.patch_size * self.merge_size, min_pixels=min_pixels,
                                        max_pixels=max_pixels)
        done = []
        for f in frames:
            if do_resize:
                f = self._resize(f, out_h, out_w, resample, input_data_format, device)
            if do_rescale:
                f = self.rescale(f, scale=rescale_factor, input_data_format=input_data_format)
            if do_normalize:
                f = self.normalize(image=f, mean=image_mean, std=image_std, input_data_format=input_data_format)
            done.append(to_channel_dimension_format(f, data_format, input_channel_dim=input_data_format))
        x = np.array(done)
        if data_format == ChannelDimension.LAST:
            x = x.transpose(0, 3, 1, 2)
        tp, ps, ms = self.temporal_patch_size, self.patch_size, self.merge_size
        if x.shape[0] == 1:  # a still image fills a whole temporal patch with copies of itself
            x = np.tile(x, (tp, 1, 1, 1))
        ch = x.shape[1]
        gt, gh, gw = x.shape[0] // tp, out_h // ps, out_w // ps
        # [t, tp, c, h/ms, ms, ps, w/ms, ms, ps] -> rows ordered (t, h/ms, w/ms, ms_h, ms_w), columns (c, tp, ps_h, ps_w)
        x = x.reshape(gt, tp, ch, gh // ms, ms, ps, gw // ms, ms, ps).transpose(0, 3, 6, 4, 7, 2, 1, 5, 8)
        return x.reshape(gt * gh * gw, ch * tp * ps * ps), (gt, gh, gw)

    @staticmethod
    def _resize(frame: np.ndarray, out_h: int, out_w: int, resample, input_data_format, device):
        if device is not None and frame.dtype == np.uint8 and resample == PILImageResampling.BICUBIC and \
                input_data_format == ChannelDimension.LAST and frame.shape[-1] == 3:
            import torch

            from . import ops

            t = torch.from_numpy(np.ascontiguousarray(frame)).to(device)
            return ops.image_resize_u8(t, (out_h, out_w))[0].cpu().numpy()  # Pillow's bicubic, bit for bit, on the GPU
        return resize(frame, size=(out_h, out_w), resample=resample, input_data_format=input_data_format)

    def preprocess(self, images, videos=None, do_resize=None, size=None, resample=None, do_rescale=None,
                   rescale_factor=None, do_normalize=None, image_mean=None, image_std=None, do_convert_rgb=None,
                   return_tensors=None, data_format=ChannelDimension.FIRST, input_data_format=None, device=None):
        """image_processing_bailingmm.py:317-462: `images` -> {"pixel_values", "image_grid_thw"}; `videos` ->
        {"pixel_values_videos", "video_grid_thw"} (videos take their own pixel budget)."""
        pick = lambda v, d: d if v is None else v  # noqa: E731
        kw = dict(do_resize=pick(do_resize, self.do_resize), resample=pick(resample, self.resample),
                  do_rescale=pick(do_rescale, self.do_rescale), rescale_factor=pick(rescale_factor, self.rescale_factor),
                  do_normalize=pick(do_normalize, self.do_normalize), image_mean=pick(image_mean, self.image_mean),
                  image_std=pick(image_std, self.image_std), do_convert_rgb=pick(do_convert_rgb, self.do_convert_rgb),
                  data_format=data_format, input_data_format=input_data_format, device=device)
        if kw["do_rescale"] and kw["rescale_factor"] is None:
            raise ValueError("`rescale_factor` must be specified if `do_rescale` is `True`.")
        if kw["do_normalize"] and (kw["image_mean"] is None or kw["image_std"] is None):
            raise ValueError("`image_mean` and `image_std` must both be specified if `do_normalize` is `True`.")
        data = {}
        for key, grid_key, samples, lo, hi in (
                ("pixel_values", "image_grid_thw", None if images is None else _as_image_batches(images),
                 self.min_pixels, self.max_pixels),
                ("pixel_values_videos", "video_grid_thw", None if videos is None else _as_video_batches(videos),
                 self.min_pixels_video, self.max_pixels_video)):
            if samples is None:
                continue
            if key == "pixel_values" and not all(is_valid_image(s) for s in samples):
                raise ValueError("Invalid image type. Must be of type PIL.Image.Image, numpy.ndarray, torch.Tensor, "
                                 "tf.Tensor or jax.ndarray.")
            rows, grids = [], []
            for sample in samples:
                patches, grid = self._preprocess(sample, min_pixels=lo, max_pixels=hi, **kw)
                rows.extend(patches)
                grids.append(grid)
            data = {key: np.array(rows), grid_key: np.array(grids)}  # (videos replace images, as in the reference)
        return BatchFeature(data=data, tensor_type=return_tensors)
