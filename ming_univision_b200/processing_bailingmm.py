"""`BailingMMProcessor` — the host side of a request in front of `MingUniVisionForConditionalGeneration.generate`
(reference: mingunivision/processing_bailingmm.py, bailingmm_utils.py:94-164, 472-537): chat template, image fetching,
the two image transforms, `<IMAGE>` placeholder expansion, tokenisation, and the two classifier-free-guidance masks the
generation loop turns into its CFG rows.

    MingTokUndProcessor(image_size=224, mean=None, std=None)         processing_bailingmm.py:80-100
        Resize((S, S), BICUBIC) -> ToTensor -> Normalize             (understanding: S = 1024, :175)
    MingTokCenterCropProcessor(image_size=224, mean=None, std=None)  processing_bailingmm.py:102-123
        Resize(S, BICUBIC) -> CenterCrop(S) -> ToTensor -> Normalize (generation / editing: S = 512, :176)
    BailingMMProcessor(image_processor, tokenizer, chat_template)    processing_bailingmm.py:125-529
        .apply_chat_template(conversation, system_template=None, add_generation_prompt=True) -> str        :377-437
        .process_vision_info(conversations) -> (images | None, None, None)                    bailingmm_utils.py:503-537
        .__call__(images=None, text=None, for_edit=False, image_patch_size=32, ...) -> BatchFeature       :179-280
              input_ids, attention_mask, uncond_attention_mask, text_uncond_attention_mask [, pixel_values, image_grid_thw]
        .tokenize(text) / .batch_decode / .decode
    load_tokenizer(dir): the reference's `tokenizer.json` (+ tokenizer_config.json) through `PreTrainedTokenizerFast`

The image transforms are the device kernels of `mingtok.utils.processor` (one C-ABI call, bit-exact with the PIL /
torchvision stack).  Everything else here is string / integer work on the host; its parity bar is EXACT and it is tested
against the live reference class (`tests/test_processing_cpu.py`: same strings, same token ids, same masks).  Why it is a
module of its own and not the reference's file imported from a checkout: the reference's processor and `BailingTokenizer`
are written against transformers 4.52 (`ProcessorMixin` / tokenizer internals) and no longer construct under the
transformers 5 of this image; the vocabulary itself is DATA (`tokenizer.json`) and loads with the stock fast tokenizer.
Videos and audio are other products' inputs and are refused.
"""
from __future__ import annotations

import base64
import io
import json
import math
import os
from typing import Dict, List, Optional, Sequence, Union

import torch

from .mingtok.utils.processor import _CLIP_MEAN, _CLIP_STD, CenterCropProcessor, ResizeProcessor

# special tokens of the chat / image markup (processing_bailingmm.py:41-67)
DEFAULT_IMAGE_PATCH_TOKEN = "<imagePatch>"
DEFAULT_IM_START_TOKEN = "<image>"
DEFAULT_IM_END_TOKEN = "</image>"
USER_PREFIX = "<role>HUMAN</role>"
ASSISTANT_PREFIX = "<role>ASSISTANT</role>"
END_OF_TEXT = "<|endoftext|>"
IMAGE_PLACEHOLDER = "<IMAGE>"

# bailingmm_utils.py:29-32
IMAGE_FACTOR = 28
MIN_PIXELS = 4 * 28 * 28
MAX_PIXELS = 1024 * 28 * 28
MAX_RATIO = 200


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


# ---------------------------------------------------------------------------------------------------------------
# image fetching (bailingmm_utils.py:94-164)
# ---------------------------------------------------------------------------------------------------------------
def fetch_size(height: int, width: int, factor: int = IMAGE_FACTOR, min_pixels: int = MIN_PIXELS,
               max_pixels: int = MAX_PIXELS) -> tuple:
    """bailingmm_utils.py:94-120 (`smart_resize` of the FETCH step; not the image processor's variant, which refuses sides
    below the factor): sides rounded to multiples of `factor` (at least one factor), area inside the pixel budget."""
    if max(height, width) / min(height, width) > MAX_RATIO:
        raise ValueError(f"absolute aspect ratio must be smaller than {MAX_RATIO}, "
                         f"got {max(height, width) / min(height, width)}")
    mult = lambda v, fn: fn(v / factor) * factor  # noqa: E731
    h, w = max(factor, mult(height, round)), max(factor, mult(width, round))
    if h * w > max_pixels:
        shrink = math.sqrt((height * width) / max_pixels)
        h, w = mult(height / shrink, math.floor), mult(width / shrink, math.floor)
    elif h * w < min_pixels:
        grow = math.sqrt(min_pixels / (height * width))
        h, w = mult(height * grow, math.ceil), mult(width * grow, math.ceil)
    return h, w


def fetch_image(ele: dict, size_factor: int = IMAGE_FACTOR):
    """bailingmm_utils.py:122-164: PIL image / http(s) url / file:// / data:image;base64 / local path -> RGB PIL image,
    resized (Pillow's default bicubic) to the `fetch_size` of its own size or of `resized_height` x `resized_width`."""
    from PIL import Image

    src = ele["image"] if "image" in ele else ele["image_url"]
    img = None
    if isinstance(src, Image.Image):
        img = src
    elif src.startswith(("http://", "https://")):
        import requests

        img = Image.open(requests.get(src, stream=True).raw)
    elif src.startswith("file://"):
        img = Image.open(src[len("file://"):])
    elif src.startswith("data:image"):
        if "base64," in src:
            img = Image.open(io.BytesIO(base64.b64decode(src.split("base64,", 1)[1])))
    else:
        img = Image.open(src)
    if img is None:
        raise ValueError(f"Unrecognized image input, support local path, http url, base64 and PIL.Image, got {src}")
    img = img.convert("RGB")
    if "resized_height" in ele and "resized_width" in ele:
        h, w = fetch_size(ele["resized_height"], ele["resized_width"], factor=size_factor)
    else:
        h, w = fetch_size(img.height, img.width, factor=size_factor, min_pixels=ele.get("min_pixels", MIN_PIXELS),
                          max_pixels=ele.get("max_pixels", MAX_PIXELS))
    return img.resize((w, h))


def process_vision_info(conversations) -> tuple:
    """bailingmm_utils.py:472-537 for image inputs: every content element carrying an image (or a list of images), in
    conversation order -> (list of fetched PIL images | None, None, None).  Video / audio elements are refused."""
    if isinstance(conversations[0], dict):
        conversations = [conversations]
    images = []
    for conversation in conversations:
        for message in conversation:
            if not isinstance(message["content"], list):
                continue
            for ele in message["content"]:
                if "video" in ele or "audio" in ele or ele.get("type") == "video":
                    raise NotImplementedError("video / audio inputs are outside the Ming-UniVision image path")
                if not ("image" in ele or "image_url" in ele or ele.get("type") in ("image", "image_url")):
                    continue
                if "image" in ele and isinstance(ele["image"], (tuple, list)):
                    images += [fetch_image({"type": "image", "image": one}) for one in ele["image"]]
                elif "image" in ele or "image_url" in ele:
                    images.append(fetch_image(ele))
                else:
                    raise ValueError("image, image_url, video, video_url, audio or audio_url should in content.")
    return (images or None), None, None


# ---------------------------------------------------------------------------------------------------------------
# tokenizer (data files of the checkpoint directory)
# ---------------------------------------------------------------------------------------------------------------
def load_tokenizer(directory: str):
    """The checkpoint's `tokenizer.json` through the stock `PreTrainedTokenizerFast`, with the special-token settings of
    `tokenizer_config.json` (bos / eos / pad / cls, the additional special tokens, no bos / eos insertion, no clean-up of
    tokenisation spaces — tokenizer_config.json of the reference).  The reference's `BailingTokenizer` subclass
    (tokenization_bailing.py) adds chat-format helpers on top of the same vocabulary; ids are identical."""
    from transformers import PreTrainedTokenizerFast

    path = os.path.join(directory, "tokenizer.json")
    if not os.path.isfile(path):
        raise FileNotFoundError(f"{path}: the tokenizer data of the checkpoint is needed")
    cfg = {}
    cfg_path = os.path.join(directory, "tokenizer_config.json")
    if os.path.isfile(cfg_path):
        with open(cfg_path) as f:
            cfg = json.load(f)
    kw = {k: cfg[k] for k in ("bos_token", "eos_token", "pad_token", "cls_token", "additional_special_tokens") if k in cfg}
    return PreTrainedTokenizerFast(tokenizer_file=path, clean_up_tokenization_spaces=False, **kw)


def _find_all(seq: Sequence[int], sub: Sequence[int]) -> List[int]:
    """Start positions of every occurrence of `sub` in `seq` (:364-372)."""
    m = len(sub)
    if m == 0:
        return []
    sub = list(sub)
    return [i for i in range(len(seq) - m + 1) if list(seq[i:i + m]) == sub]


def cfg_masks(seq: Sequence[int], user_prefix_ids: Sequence[int], assistant_prefix_ids: Sequence[int],
              image_token_ids) -> tuple:
    """The two classifier-free-guidance masks of one token sequence (processing_bailingmm.py:305-351), as 0 / 1 lists.

    Both concern the LAST `<role>HUMAN</role>` turn: its body runs from behind that tag to the first
    `<role>ASSISTANT</role>` tag at or after it.
      uncond       hides the whole body (text AND image) — but only when that ASSISTANT tag exists; an open turn keeps
                   the mask all ones;
      text_uncond  hides the body's TEXT and keeps its image tokens (`<image>`, `<imagePatch>`, `</image>`); an open turn
                   runs to the end of the sequence."""
    n = len(seq)
    uncond, text_uncond = [1] * n, [1] * n
    users = _find_all(seq, user_prefix_ids)
    if not users:
        return uncond, text_uncond
    last_user = users[-1]
    closing = next((p for p in _find_all(seq, assistant_prefix_ids) if p >= last_user), None)
    body = last_user + len(user_prefix_ids)
    if closing is not None:
        for i in range(body, closing):
            uncond[i] = 0
    for i in range(body, n if closing is None else closing):
        if seq[i] not in image_token_ids:
            text_uncond[i] = 0
    return uncond, text_uncond


class BailingMMProcessor:
    """See the module docstring.  `vis_processor` / `gen_processor` default to the device transforms (1024 resize for
    understanding, 512 centre crop for generation / editing, mean = std = 0.5: processing_bailingmm.py:175-176); pass
    replacements to run elsewhere (the CPU tests pass the torchvision stacks)."""

    attributes = ["image_processor", "tokenizer"]

    def __init__(self, image_processor=None, tokenizer=None, chat_template=None, image_token="<image>",
                 video_token="<video>", audio_token="<audio>", vis_processor=None, gen_processor=None, device="cuda",
                 **kwargs):
        if tokenizer is None:
            raise ValueError("BailingMMProcessor needs a tokenizer")
        self.image_processor, self.tokenizer = image_processor, tokenizer
        self.image_token, self.video_token, self.audio_token = image_token, video_token, audio_token
        half = [0.5, 0.5, 0.5]
        self.vis_processor = vis_processor if vis_processor is not None else MingTokUndProcessor(
            image_size=1024, mean=half, std=half, device=device)
        self.gen_processor = gen_processor if gen_processor is not None else MingTokCenterCropProcessor(
            image_size=512, mean=half, std=half, device=device)
        self.chat_template = chat_template if chat_template is not None else getattr(tokenizer, "chat_template", None)
        self.gen_terminator = [tokenizer.convert_tokens_to_ids(END_OF_TEXT)]

    @classmethod
    def from_pretrained(cls, directory: str, **kwargs):
        """Tokenizer from the directory's data files; the video image processor from its `preprocessor_config.json`."""
        from .image_processing_bailingmm import BailingMMImageProcessor

        ip_kw = {}
        cfg_path = os.path.join(directory, "preprocessor_config.json")
        if os.path.isfile(cfg_path):
            with open(cfg_path) as f:
                cfg = json.load(f)
            ip_kw = {k: cfg[k] for k in ("min_pixels", "max_pixels", "patch_size", "temporal_patch_size", "merge_size",
                                         "image_mean", "image_std") if k in cfg}
        return cls(image_processor=BailingMMImageProcessor(**ip_kw), tokenizer=load_tokenizer(directory), **kwargs)

    # ---- strings ------------------------------------------------------------------------------------------------
    def apply_system_template(self, text: str) -> str:
        return USER_PREFIX

    def apply_chat_template(self, conversation: List[Dict], system_template: Optional[str] = None, **kwargs) -> str:
        """:377-437.  A conversation is a list of {"role": "HUMAN" | "ASSISTANT", "content": [elements]}; the system part is
        the opening HUMAN tag, ASSISTANT turns are closed with `<|endoftext|>` + the next HUMAN tag, image elements leave
        `<IMAGE>` placeholders (one per image that the message's text does not already mark with `<image>`), and the string
        ends with the ASSISTANT tag unless `add_generation_prompt=False`.  `tokenize=` / `use_system=` are accepted and
        ignored, as there."""
        from PIL import Image

        out = []
        for message in conversation:
            role = message["role"]
            if role not in ("HUMAN", "ASSISTANT"):
                raise AssertionError(f"role {role!r}: HUMAN or ASSISTANT")
            if role == "ASSISTANT":
                out.append(ASSISTANT_PREFIX)
            marked = str(message["content"]).count("<image>")  # (counted on the repr of the whole content list, as there)
            for ele in message["content"]:
                kind = ele["type"]
                if kind == "image":
                    n = 1 if isinstance(ele["image"], (str, Image.Image)) else len(ele["image"])
                    if marked < n:
                        out.append("\n".join([IMAGE_PLACEHOLDER] * (n - marked)))
                elif kind == "text":
                    out.append(ele["text"])
                elif kind in ("video", "audio"):
                    raise NotImplementedError("video / audio inputs are outside the Ming-UniVision image path")
            if role == "ASSISTANT":
                out.append(END_OF_TEXT + USER_PREFIX)
        if kwargs.get("add_generation_prompt", True):
            out.append(ASSISTANT_PREFIX)
        text = "".join(out)
        head = system_template if system_template is not None else self.apply_system_template(text)
        return head + text

    def process_vision_info(self, conversations):
        return process_vision_info(conversations)

    def _expand_image_tokens(self, text: List[str], image_grid_thw, special_token: str = IMAGE_PLACEHOLDER) -> List[str]:
        """:445-464: the i-th `<IMAGE>` placeholder (counted over all samples) becomes
        `<image>` + t*h*w x `<imagePatch>` + `</image>` + newline."""
        counts = [int(v) for v in torch.as_tensor(image_grid_thw).prod(dim=1).tolist()]
        out, used = [], 0
        for sample in text:
            for _ in range(sample.count(special_token)):
                block = DEFAULT_IM_START_TOKEN + DEFAULT_IMAGE_PATCH_TOKEN * counts[used] + DEFAULT_IM_END_TOKEN + "\n"
                sample = sample.replace(special_token, block, 1)
                used += 1
            out.append(sample)
        return out

    # ---- ids and masks ------------------------------------------------------------------------------------------
    def tokenize(self, text, **output_kwargs) -> dict:
        """:282-361: token ids + attention mask from the tokenizer, and the two CFG masks (`cfg_masks`).  As in the
        reference, only a nested `text_kwargs={...}` reaches the tokenizer: the flat keyword arguments `__call__` hands
        over (`return_tensors`, `padding`) are looked up under that key and therefore dropped."""
        enc = self.tokenizer(text, **output_kwargs.get("text_kwargs", {}))
        ids, am = enc["input_ids"], enc["attention_mask"]
        if isinstance(ids, (list, tuple)) and not isinstance(ids[0], (list, tuple)):
            ids, am = [ids], ([am] if am is not None else None)
        user = self.tokenizer.encode(USER_PREFIX, add_special_tokens=False)
        assistant = self.tokenizer.encode(ASSISTANT_PREFIX, add_special_tokens=False)
        image_ids = set(self.tokenizer.convert_tokens_to_ids([DEFAULT_IM_START_TOKEN, DEFAULT_IMAGE_PATCH_TOKEN,
                                                              DEFAULT_IM_END_TOKEN]))
        masks = [cfg_masks(seq, user, assistant, image_ids) for seq in ids]
        as_long = lambda rows: torch.tensor(rows, dtype=torch.long)  # noqa: E731
        return {"input_ids": as_long(ids), "attention_mask": as_long(am) if am is not None else None,
                "uncond_attention_mask": as_long([m[0] for m in masks]),
                "text_uncond_attention_mask": as_long([m[1] for m in masks])}

    def __call__(self, images=None, videos=None, audios=None, text: Union[str, List[str], None] = None,
                 for_edit: Optional[bool] = False, **kwargs):
        """:179-280: per image the understanding transform (`for_edit=False`: 1024 x 1024) or the generation / editing
        transform (`for_edit=True`: 512 centre crop) and its `[1, H / patch, W / patch]` grid; placeholder expansion;
        `tokenize`.  Returns a `BatchFeature`; `pixel_values` stays on the device the transforms put it on."""
        from transformers.feature_extraction_utils import BatchFeature

        if videos is not None or audios is not None:
            raise NotImplementedError("video / audio inputs are outside the Ming-UniVision image path")
        patch = kwargs.pop("image_patch_size", 32)
        if isinstance(text, str):
            text = [text]
        elif not isinstance(text, list) or (text and not isinstance(text[0], str)):
            raise ValueError("Invalid input text. Please provide a string, or a list of strings")
        data = {}
        if images is not None:
            transform = self.gen_processor if for_edit else self.vis_processor
            pixels = [transform(img) for img in images]
            grid = torch.tensor([[1, p.shape[1] // patch, p.shape[2] // patch] for p in pixels], dtype=torch.long)
            data = {"pixel_values": torch.stack(pixels), "image_grid_thw": grid}
            text = self._expand_image_tokens(text, grid)
        kwargs.pop("padding_side", None)
        return BatchFeature(data={**self.tokenize(text, **kwargs), **data})

    def batch_decode(self, *args, **kwargs):
        return self.tokenizer.batch_decode(*args, **kwargs)

    def decode(self, *args, **kwargs):
        return self.tokenizer.decode(*args, **kwargs)

    @property
    def model_input_names(self):
        names = list(getattr(self.tokenizer, "model_input_names", []))
        names += list(getattr(self.image_processor, "model_input_names", []) or [])
        return list(dict.fromkeys(names))
