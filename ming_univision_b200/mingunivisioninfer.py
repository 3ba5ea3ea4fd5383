"""Inference façade with the reference's surface (mingunivision/mingunivisioninfer.py:28-120):

    MingUniVisionInfer(model_name_or_path, dtype="bf16").generate(messages, max_new_tokens=512,
                                                                  output_image_prefix="output", for_edit=False) -> str
    .reset_inner_state()

The chat template, image fetching, placeholder expansion, tokenisation and the CFG masks are this package's
`processing_bailingmm.BailingMMProcessor` (exact mirror of the reference's host code, tests/test_processing_cpu.py); only
the tokenizer DATA (`tokenizer.json`, `tokenizer_config.json`, `preprocessor_config.json`) is read from `reference_dir`
(default: the reference's "./mingunivision", mingunivisioninfer.py:43-44) or, if it holds them, from the checkpoint
directory.  Everything behind `self.model.generate(...)` is the B200-native path of this repo.
`dtype="int4" / "int8"` (bitsandbytes / quanto) are out of scope and raise.

Tests inject `model`, `processor` and `tokenizer` directly; no pretrained checkpoint exists offline, so the checkpoint
loader below (`load_checkpoint`: config.json + *.safetensors of the HF layout, `mingunivisioninfer.py:72-78`) is exercised
on synthetic state dicts only.
"""
from __future__ import annotations

import glob
import json
import os

import torch

from .mingtok.modeling_mingtok import MingTokConfig
from .modeling_bailing_moe import BailingMoeConfig
from .modeling_bailingmm import MingUniVisionForConditionalGeneration


def _mingtok_dir(model_dir: str) -> str:
    """`models/MingTok-Vision` inside the checkpoint directory, else relative to the working directory — where the
    reference looks (`MingTok.from_pretrained("./models/MingTok-Vision")`, modeling_bailingmm.py:102)."""
    inside = os.path.join(model_dir, "models", "MingTok-Vision")
    return inside if os.path.isdir(inside) else os.path.join(".", "models", "MingTok-Vision")


def resolve_checkpoint_dir(name_or_path: str) -> str:
    """A local directory is used as it is; anything else is taken for a Hugging Face hub id (the reference passes
    "inclusionAI/Ming-UniVision-16B-A3B" to `from_pretrained`, README) and fetched / found in the local cache with
    `huggingface_hub.snapshot_download`."""
    if os.path.isdir(name_or_path):
        return name_or_path
    try:
        from huggingface_hub import snapshot_download
    except ImportError as e:  # pragma: no cover
        raise FileNotFoundError(f"{name_or_path} is not a directory and huggingface_hub is not installed") from e
    return snapshot_download(name_or_path)


def build_from_config(model_dir: str, device="cuda", ep_rank: int = 0,
                      ep_size: int = 1) -> MingUniVisionForConditionalGeneration:
    """The wrapper a checkpoint directory's `config.json` describes, with uninitialised bf16 storage on `device` (`"meta"`:
    no storage at all — the full 16B-A3B parameter tree can be inspected anywhere)."""
    with open(os.path.join(model_dir, "config.json")) as f:
        cfg = json.load(f)
    tok_cfg = cfg.get("mingtok_config")
    if tok_cfg is None:
        with open(os.path.join(_mingtok_dir(model_dir), "config.json")) as f:
            tok_cfg = json.load(f)
    llm_cfg = {k: v for k, v in cfg["llm_config"].items() if k not in ("architectures", "model_type", "torch_dtype",
                                                                       "transformers_version", "auto_map")}
    # The reference's own callers build 2-D position ids, so its rotary embedding is the 1-D legacy one whatever the
    # checkpoint's `rope_scaling` says about "3D"  — mapped to None, loudly.  Any OTHER scaling (yarn,
    # linear, dynamic NTK) changes the arithmetic and is not implemented: refuse rather than mis-execute.
    rs = llm_cfg.get("rope_scaling")
    if rs is not None:
        kind = rs.get("type", rs.get("rope_type")) if isinstance(rs, dict) else rs
        if kind not in ("3D", "default"):
            raise NotImplementedError(f"rope_scaling {rs!r}: only None or the reference's '3D' entry are supported")
        import warnings

        warnings.warn("config rope_scaling '3D' -> None: the generation path uses 2-D position ids and the 1-D legacy "
                      "rotary embedding (pass 3-D position ids to forward_tokens for the M-RoPE variant)")
    llm_cfg["rope_scaling"] = None
    if cfg.get("mlp_depth", 2) != 2:
        raise NotImplementedError(f"mlp_depth {cfg['mlp_depth']}: linear_proj is Linear-GELU-Linear on this path "
                                  "(mingunivision/config.json:120)")
    # the repository's own config.json carries no `vishead_diffloss_config`: the defaults of setup_vishead_diffloss
    # (width 3072, depth 12, 16 steps, flow_matching_swiglu-4; modeling_bailing_moe.py:1559-1567) are the released head
    vishead_cfg = dict(cfg.get("vishead_diffloss_config") or {})
    if not (ep_size >= 1 and 0 <= ep_rank < ep_size):
        raise ValueError(f"ep_rank {ep_rank} / ep_size {ep_size}")
    model = MingUniVisionForConditionalGeneration.on_device(BailingMoeConfig(**llm_cfg), MingTokConfig(**tok_cfg),
                                                            vishead_cfg, device, ep_rank=ep_rank, ep_size=ep_size)
    return model


_LOAD_CHUNK_BYTES = 2 << 30  # staging memory of load_checkpoint: tensors are copied into the model in chunks of this size


def load_checkpoint(model_dir: str, device="cuda", ep_rank: int = 0,
                    ep_size: int = 1) -> MingUniVisionForConditionalGeneration:
    """Builds the wrapper from an HF-layout checkpoint directory: `config.json` with `llm_config`,
    `vishead_diffloss_config` (modeling_bailingmm.py:93-129) and the MingTok config either inline (`mingtok_config`) or
    in `models/MingTok-Vision/config.json` (the reference's hard-wired relative path, :102); weights from every
    `*.safetensors` shard (reference key schema).

    Direct safetensors -> kernel layout: the module tree is built on the meta device with bf16 storage
    (no HF `from_pretrained`, no fp32 copy, no weight init), the tensors are read straight from the
    memory-mapped shards (`safe_open`, 2 GiB of staging at a time) and copied into place — the routed experts directly into the two contiguous slabs
    per layer that the expert kernels stream, so they are never re-packed.  With `ep_size` > 1 (expert parallelism) only
    this rank's 64 / ep_size routed experts are allocated AND read: the other ranks' expert tensors are skipped without
    touching their bytes (follow with `model.model.model.set_expert_parallel(...)`)."""
    from safetensors import safe_open

    model_dir = resolve_checkpoint_dir(model_dir)
    model = build_from_config(model_dir, device, ep_rank, ep_size)
    state = model.state_dict()
    expected = {k for k, v in state.items() if not v.is_meta}
    elsewhere = {k for k, v in state.items() if v.is_meta}  # routed experts owned by other ranks
    del state
    loaded = set()
    shards = [(s, "") for s in sorted(glob.glob(os.path.join(model_dir, "*.safetensors")))]
    tok_dir = _mingtok_dir(model_dir)
    shards += [(s, "vision.") for s in sorted(glob.glob(os.path.join(tok_dir, "*.safetensors")))]
    unexpected = []
    part, part_bytes = {}, 0

    def flush():  # (one load_state_dict walks the whole module tree: batch the tensors, bounded staging memory)
        nonlocal part, part_bytes
        if part:
            model.load_state_dict(part, strict=False)
            loaded.update(part)
        part, part_bytes = {}, 0

    for shard, prefix in shards:
        with safe_open(shard, framework="pt", device=str(device)) as f:
            for key in f.keys():
                k = prefix + key
                # only the three sub-trees of the continuous-visual-token path (audio encoder / talker weights of an
                # omni checkpoint are ignored); the reference-only rotary buffers are derived from rope_theta here
                if k.split(".")[0] not in ("model", "vision", "linear_proj") or k.endswith("rotary_emb.inv_freq"):
                    continue
                if k in elsewhere:
                    continue
                if k not in expected:
                    unexpected.append(k)
                    continue
                t = f.get_tensor(key)
                part[k] = t
                part_bytes += t.numel() * t.element_size()
                if part_bytes >= _LOAD_CHUNK_BYTES:
                    flush()
            flush()
    missing = sorted(expected - loaded)
    if missing or unexpected:
        raise RuntimeError(f"checkpoint does not match the model: missing {missing[:5]}, unexpected {unexpected[:5]}")
    return model


class MingUniVisionInfer:
    def __init__(self, model_name_or_path, dtype="bf16", *, model=None, processor=None, tokenizer=None,
                 reference_dir="./mingunivision"):
        if dtype != "bf16":
            raise NotImplementedError("int4 / int8 loading (bitsandbytes / quanto) is outside the B200-native path")
        self.model_name_or_path = model_name_or_path
        self.dtype = dtype
        model_dir = str(model_name_or_path) if model is not None else resolve_checkpoint_dir(str(model_name_or_path))
        if processor is None or tokenizer is None:
            tokenizer, processor = self._load_processor(model_dir, reference_dir, tokenizer, processor)
        self.tokenizer, self.processor = tokenizer, processor
        self.model = model if model is not None else load_checkpoint(model_dir)
        self.model.tokenizer = self.tokenizer
        self.model.model.tokenizer = self.tokenizer

    @staticmethod
    def _load_processor(model_dir, reference_dir: str, tokenizer=None, processor=None):
        """mingunivisioninfer.py:43-44 (`AutoTokenizer` / `AutoProcessor.from_pretrained("./mingunivision")`): the tokenizer
        data files from the checkpoint directory if it has them, else from `reference_dir`; the processor is this
        package's mirror (the reference's own classes target transformers 4.52 and do not construct under 5.x)."""
        from .processing_bailingmm import BailingMMProcessor, load_tokenizer

        data_dir = next((d for d in (str(model_dir), os.path.abspath(reference_dir))  # (a hub id is not a directory: skipped)
                         if os.path.isfile(os.path.join(d, "tokenizer.json"))), None)
        if tokenizer is None:
            if data_dir is None:
                raise RuntimeError(f"tokenizer.json found neither in {model_dir} nor in {os.path.abspath(reference_dir)} "
                                   "(pass processor= and tokenizer= to inject your own)")
            tokenizer = load_tokenizer(data_dir)
        if processor is None:
            processor = (BailingMMProcessor.from_pretrained(data_dir) if data_dir is not None
                         else BailingMMProcessor(tokenizer=tokenizer))
            processor.tokenizer = tokenizer
        return tokenizer, processor

    @property
    def device(self):
        return next(self.model.parameters()).device

    @torch.no_grad()
    def generate(self, messages, max_new_tokens=512, output_image_prefix="output", for_edit=False):
        """mingunivisioninfer.py:82-117, call for call."""
        text = self.processor.apply_chat_template(messages, tokenize=False, add_generation_prompt=True, use_system=True)
        image_inputs, _, _ = self.processor.process_vision_info(messages)
        inputs = self.processor(text=[text], images=image_inputs, return_tensors="pt",
                                image_patch_size=self.model.vision.patch_size, for_edit=for_edit).to(self.device)
        kw = dict(inputs)
        if kw.get("pixel_values") is not None:
            kw["pixel_values"] = kw["pixel_values"].to(dtype=torch.bfloat16)
        generated_ids = self.model.generate(**kw, max_new_tokens=max_new_tokens, use_cache=True,
                                            output_image_prefix=output_image_prefix)
        trimmed = [out_ids[len(in_ids):] for in_ids, out_ids in zip(inputs["input_ids"], generated_ids)]
        return self.processor.batch_decode(trimmed, skip_special_tokens=True, clean_up_tokenization_spaces=False)[0]

    def reset_inner_state(self):
        self.model.reset_inner_state()
