"""B200-native multimodal wrapper: the continuous-visual-token surface of the reference's
``mingunivision/modeling_bailingmm.py`` (MingUniVisionForConditionalGeneration :85-308) — `vision` (MingTok), `model`
(BailingMoeForCausalLM with vis_head + diffloss), `linear_proj`, `extract_image_feature`, `prompt_wrap_vision`,
`reset_inner_state`, and `generate` with its multi-round state (KV cache + three attention masks carried across calls,
:206-301).  HF GenerationMixin is not used (its transformers-5 internals no longer match the reference's 4.52 code,
SURVEY.md §7): `generate` runs prefill, CUDA-graphed greedy decoding and — when the model emits `<image>` —
`generate_image` itself; `generate_text` / `generate_image_from_prompt` are the single-purpose entries.
"""
from __future__ import annotations

import os
from typing import Optional

import torch
import torch.nn as nn

from . import _packs, ops
from .mingtok.modeling_mingtok import MingTok, MingTokConfig
from .modeling_bailing_moe import BailingMoeConfig, BailingMoeForCausalLM

BF16 = torch.bfloat16


class LinearProj(nn.Sequential):
    """nn.Sequential(Linear(feature_dim, hidden), GELU, Linear(hidden, hidden)) — modeling_bailingmm.py:111-115 —
    keys linear_proj.0.* / linear_proj.2.*; executed as two GEMMs with the exact-erf GELU fused into the first."""

    def __init__(self, feature_dim: int, hidden: int):
        super().__init__(nn.Linear(feature_dim, hidden), nn.GELU(), nn.Linear(hidden, hidden))
        self._pk = None
        _packs.watch(self, self._reset_packs)

    def _reset_packs(self) -> None:
        self._pk = None

    def _apply(self, fn, *a, **k):
        self._reset_packs()
        _packs.bump()
        return super()._apply(fn, *a, **k)

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        dev = self[0].weight.device
        if dev.type != "cuda":
            raise RuntimeError("linear_proj (B200-native) runs on CUDA only; there is no CPU fallback")
        if self._pk is None or self._pk[0].device != dev:
            d = lambda t: t.detach().to(device=dev, dtype=BF16).contiguous()  # noqa: E731
            self._pk = (d(self[0].weight), d(self[0].bias), d(self[2].weight), d(self[2].bias))
        w0, b0, w2, b2 = self._pk
        x2 = x.reshape(-1, x.shape[-1])
        x2 = x2.contiguous() if x2.dtype == BF16 else ops.affine(x2, 1.0, 0.0)
        if x2.shape[0] <= 8:
            y = ops.gemv(ops.gemv(x2, w0, b0, epi=ops.EPI_GELU), w2, b2)
        else:
            y = ops.linear(ops.linear(x2, w0, b0, epi=ops.EPI_GELU), w2, b2)
        return y.view(*x.shape[:-1], -1)


class MingUniVisionForConditionalGeneration(nn.Module):
    def __init__(self, llm_config: BailingMoeConfig, mingtok_config: MingTokConfig, vishead_diffloss_config: dict,
                 mlp_depth: int = 2):
        super().__init__()
        if mlp_depth != 2:
            raise NotImplementedError("mlp_depth = 2 on the Ming-UniVision path (config.json)")
        self.vision = MingTok(mingtok_config)
        self.model = BailingMoeForCausalLM(llm_config)
        self.linear_proj = LinearProj(self.vision.feature_dim, llm_config.hidden_size)
        vh = dict(vishead_diffloss_config, hidden_size=llm_config.hidden_size,
                  image_emb_dim_for_gen=self.vision.latent_dim)
        self.model.setup_vishead_diffloss(**vh)
        self.past_key_values = None
        self.past_attention_mask = None
        self.past_text_uncond_attention_mask = None
        self.past_uncond_attention_mask = None
        self.generated_images = []
        for p in self.parameters():
            p.requires_grad_(False)

    @classmethod
    def on_device(cls, llm_config: BailingMoeConfig, mingtok_config: MingTokConfig, vishead_diffloss_config: dict,
                  device, ep_rank: int = 0, ep_size: int = 1, dtype=BF16) -> "MingUniVisionForConditionalGeneration":
        """Builds the wrapper WITHOUT ever holding an fp32 or duplicated copy of the 16.8 B parameters: the module tree is
        created on the meta device, then every parameter gets `dtype` storage on `device`; the routed experts of each
        layer live in the two contiguous slabs the expert kernels stream (BailingMoeSparseMoeBlock.materialize_experts),
        and with `ep_size` > 1 only this rank's E / ep_size experts are allocated (the others stay meta tensors).  The
        parameters are uninitialised: follow with `load_state_dict` (copies land in the slabs) or an in-place init."""
        prev = torch.get_default_dtype()
        torch.set_default_dtype(dtype)
        try:
            with torch.device("meta"):
                m = cls(llm_config, mingtok_config, vishead_diffloss_config)
        finally:
            torch.set_default_dtype(prev)
        expert_params = set()
        for lyr in m.model.model.layers:
            lyr.mlp.materialize_experts(device, ep_rank, ep_size, dtype)
            expert_params.update(id(p) for p in lyr.mlp.experts.parameters())
        for mod in m.modules():
            for name, p in list(mod._parameters.items()):
                if p is not None and p.is_meta and id(p) not in expert_params:
                    mod._parameters[name] = nn.Parameter(torch.empty(p.shape, dtype=dtype, device=device),
                                                         requires_grad=False)
            for name, b in list(mod._buffers.items()):
                if b is not None and b.is_meta:
                    mod._buffers[name] = torch.zeros(b.shape, dtype=b.dtype, device=device)
        return m

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, *args, device_map="cuda", torch_dtype=None,
                        quantization_config=None, ep_rank: int = 0, ep_size: int = 1, **kwargs):
        """The call the reference's facade makes (mingunivisioninfer.py:72-78: `from_pretrained(path,
        torch_dtype=torch.bfloat16, attn_implementation="flash_attention_2", trust_remote_code=True, device_map="cuda")`)
        on a LOCAL checkpoint directory: `mingunivisioninfer.load_checkpoint` (meta-device construction, tensors straight
        from the safetensors shards into the kernel slabs; `ep_rank` / `ep_size`: only this rank's routed experts).
        `attn_implementation` / `trust_remote_code` are accepted and ignored; quantised loading is refused."""
        from .mingunivisioninfer import load_checkpoint

        if quantization_config is not None:
            raise NotImplementedError("int4 / int8 loading (bitsandbytes / quanto) is outside the B200-native path")
        if torch_dtype not in (None, BF16, "bfloat16", "auto"):
            raise NotImplementedError(f"torch_dtype {torch_dtype!r}: the path computes in bf16")
        device = device_map if isinstance(device_map, (str, torch.device)) and device_map != "auto" else "cuda"
        return load_checkpoint(str(pretrained_model_name_or_path), device=device, ep_rank=ep_rank, ep_size=ep_size)

    def reset_inner_state(self):
        """modeling_bailingmm.py:303-308."""
        self.past_key_values = None
        self.past_attention_mask = None
        self.past_text_uncond_attention_mask = None
        self.past_uncond_attention_mask = None
        self.model.reset_image_gen_status()

    @torch.no_grad()
    def extract_image_feature(self, pixel_values, grid_thw=None):
        """:131-138: MingTok encoder + semantic decoder, then linear_proj."""
        feats = self.vision(pixel_values)["x_norm_patchtokens"]
        return self.linear_proj(feats)

    @torch.no_grad()
    def prompt_wrap_vision(self, input_ids, inputs_embeds, vision_embeds, image_token_id=None):
        """:152-177: scatter the image embeddings into the positions of the `<imagePatch>` token."""
        if vision_embeds is None or input_ids is None:
            return inputs_embeds, None
        if len(vision_embeds.shape) == 3:
            vision_embeds = vision_embeds.reshape(-1, vision_embeds.shape[-1])
        image_token_id = self.model.config.image_patch_token if image_token_id is None else image_token_id
        n_image_tokens = int((input_ids == image_token_id).sum())
        if n_image_tokens != vision_embeds.shape[0]:
            raise ValueError(f"Image features and image tokens do not match: tokens: {n_image_tokens}, "
                             f"features {vision_embeds.shape[0]}")
        image_mask = (input_ids == image_token_id)
        out = inputs_embeds.clone()
        out[image_mask] = vision_embeds.to(out.dtype)
        return out, image_mask

    @torch.no_grad()
    def generate(self, input_ids, attention_mask=None, uncond_attention_mask=None, text_uncond_attention_mask=None,
                 pixel_values=None, image_grid_thw=None, max_new_tokens: int = 512, eos_token_id=None,
                 output_image_prefix=None, image_gen_temperature=1.0, image_gen_text_cfg=3.0, image_gen_image_cfg=1.1,
                 max_cache_len: int = 4096, **generate_kwargs):
        """MingUniVisionForConditionalGeneration.generate (modeling_bailingmm.py:206-301) with its multi-round state:
        the KV cache and the three attention masks persist across calls (`reset_inner_state()` clears them), so a later
        call only prefills ITS prompt behind the cached context (in-context editing, test_infer_unified.py:30-56).

        One call = what HF GenerationMixin.generate does with do_sample = false on this model: prefill, then greedy
        tokens; when the model emits the `<image>` start token the next step is `generate_image` (256 visual tokens through
        the LLM + RF head + semantic decoder with CFG rows built from `uncond_attention_mask` /
        `text_uncond_attention_mask`, modeling_bailing_moe.py:1769-1796), after which text decoding resumes from the
        cond row's last hidden state; stops at `eos_token_id` or after `max_new_tokens` text tokens.  Returns the full
        token sequence [1, prompt + new]; generated images are collected in `self.generated_images` (and written to
        `{output_image_prefix}[_{i}].png` when a prefix is given, as the reference does from inside forward()).
        As in the reference, the CFG scales never reach the sampler (kwargs bug)."""
        llm, cfg = self.model, self.model.config
        dev = input_ids.device
        if input_ids.shape[0] != 1:
            raise ValueError("generation runs on one sequence (the reference asserts batch 1, modeling_bailing_moe.py:1865)")
        eos = cfg.pad_token_id if eos_token_id is None else eos_token_id
        S = input_ids.shape[1]
        ones = lambda n: torch.ones((1, n), dtype=torch.int32, device=dev)  # noqa: E731
        attention_mask = ones(S) if attention_mask is None else attention_mask.to(dev, torch.int32)
        # None masks stay None on the first round, exactly as in the reference: generate_image then runs B = 1 without
        # guidance (modeling_bailing_moe.py:1867-1891).  Later rounds concatenate them behind the saved ones (:229-234),
        # where the reference's torch.cat raises on None — so they are required there.
        if uncond_attention_mask is not None:
            uncond_attention_mask = uncond_attention_mask.to(dev, torch.int32)
        if text_uncond_attention_mask is not None:
            text_uncond_attention_mask = text_uncond_attention_mask.to(dev, torch.int32)
        if self.past_attention_mask is not None:  # :229-234
            if uncond_attention_mask is None or text_uncond_attention_mask is None:
                raise ValueError("rounds after the first need uncond_attention_mask and text_uncond_attention_mask "
                                 "(they are concatenated behind the saved masks, modeling_bailingmm.py:229-234)")
            attention_mask = torch.cat((self.past_attention_mask, attention_mask), dim=1)
            uncond_attention_mask = torch.cat((self.past_uncond_attention_mask, uncond_attention_mask), dim=1)
            text_uncond_attention_mask = torch.cat((self.past_text_uncond_attention_mask, text_uncond_attention_mask), dim=1)
        cache = self.past_key_values
        if cache is None:
            need = S + max_new_tokens + 2 * (cfg.num_image_tokens_for_gen + 1) + 8
            cache = llm.new_cache(max_len=max(need, max_cache_len))
            cache.seq_len, cache.batch = 0, 1
            llm._gen_ws, llm._txt_ws = {}, {}
        t0 = cache.seq_len
        if attention_mask.shape[1] != t0 + S:
            raise ValueError(f"attention mask length {attention_mask.shape[1]} != cached context {t0} + prompt {S}")
        if bool((attention_mask == 0).any()):
            raise NotImplementedError("the cond row's attention mask is all ones on this path (left padding is not used)")
        emb = llm.model.embed(input_ids.clamp(0, cfg.vocab_size - 1))
        image_mask = None
        if pixel_values is not None and S > 1:
            emb, image_mask = self.prompt_wrap_vision(input_ids, emb, self.extract_image_feature(pixel_values, image_grid_thw))
        pos = torch.arange(t0, t0 + S, device=dev, dtype=torch.int32).unsqueeze(0)
        hidden = llm.model.forward_tokens(emb, pos, cache, key_mask=None, image_mask=image_mask)
        last = hidden[:, -1]
        new_ids: list = []
        self.generated_images = []
        budget = max_new_tokens
        while budget > 0:
            out = llm.greedy_decode(last, cache, budget, stop_ids=(eos, cfg.image_start_token))
            new_ids += out
            budget -= len(out)
            if not out or out[-1] != cfg.image_start_token:
                break
            # the model asked for an image: `<image>` is the first input of generate_image (:1769-1796)
            am = ones(cache.seq_len + 1)
            start = llm.model.embed(torch.tensor([[cfg.image_start_token]], device=dev))
            img, hid, _ = llm.generate_image(
                input_embeds=start, past_key_values=cache, attention_mask=am,
                uncond_attention_mask=uncond_attention_mask, text_uncond_attention_mask=text_uncond_attention_mask,
                latent_to_sem_func=self.vision.forward_feature_decoder, linear_proj=self.linear_proj,
                sem_to_pix_func=self.vision.forward_pixel_decoder, image_gen_text_cfg=image_gen_text_cfg,
                image_gen_image_cfg=image_gen_image_cfg, image_gen_temperature=image_gen_temperature)
            self.generated_images.append(img[0:1])
            llm.num_generated_images += 1
            if output_image_prefix is not None:
                self._save_image(img[0], output_image_prefix, len(self.generated_images) - 1)
            last = hid[0:1, -1]  # cond row: its logits choose the token after the image (:1783-1787)
        # ---- state for the next round (:272-299)
        self.past_key_values = cache
        pad_n = cache.seq_len - attention_mask.shape[1]
        pad1 = ones(pad_n)
        pad0 = torch.zeros((1, pad_n), dtype=torch.int32, device=dev)
        past_mode = os.environ.get("PAST_MODE", "DROP")
        if past_mode == "KEEP":
            if uncond_attention_mask is None or text_uncond_attention_mask is None:
                raise ValueError("PAST_MODE=KEEP carries the two uncond masks to the next round; they cannot be None")
            self.past_attention_mask = torch.cat((attention_mask, pad1), dim=1)
            self.past_text_uncond_attention_mask = torch.cat((text_uncond_attention_mask, pad1), dim=1)
            self.past_uncond_attention_mask = torch.cat((uncond_attention_mask, pad0), dim=1)
        elif past_mode == "DROP":
            self.past_attention_mask = torch.cat((attention_mask, pad1), dim=1)
            self.past_text_uncond_attention_mask = torch.cat((attention_mask, pad1), dim=1)
            self.past_uncond_attention_mask = torch.cat((attention_mask, pad0), dim=1)
        else:
            raise ValueError("PAST_MODE must be KEEP or DROP")
        llm.reset_image_gen_status()
        llm.model.check_expert_parallel()
        return torch.cat((input_ids, torch.tensor([new_ids], dtype=input_ids.dtype, device=dev)), dim=1)

    @staticmethod
    def _save_image(img: torch.Tensor, prefix: str, index: int) -> None:
        """[-1, 1] CHW tensor -> `{prefix}.png` / `{prefix}_{i}.png` (modeling_bailing_moe.py:84-90, :1788-1796)."""
        from PIL import Image

        # tensor_to_pil (:84-90): x*std + mean in fp32, then ToPILImage = mul(255).byte(), i.e. TRUNCATION
        arr = ((img.float().clamp(-1, 1) * 0.5 + 0.5) * 255).to(torch.uint8).permute(1, 2, 0).cpu().numpy()
        Image.fromarray(arr).save(f"{prefix}.png" if index == 0 else f"{prefix}_{index}.png")

    @torch.no_grad()
    def generate_image_from_prompt(self, input_ids, pixel_values=None, uncond_attention_mask=None,
                                   text_uncond_attention_mask=None, image_gen_temperature=1.0, noises=None):
        """Prefill `input_ids` (optionally with an input image scattered at its `<imagePatch>` positions), then run
        BailingMoeForCausalLM.generate_image on the `<image>` start token — the sequence forward() performs when HF
        generate feeds it that token (modeling_bailing_moe.py:1769-1796).  Returns (image [1,3,H,W], final mask).

        Batched serving (extension): input_ids [G, S] — G requests of equal prompt length (masks [G, S+1],
        pixel_values [G, 3, H, W]) — are prefilled and generated TOGETHER, G x CFG rows <= 8; returns images [G, 3, H, W].
        Every request's result is what it would be on its own (the kernels are row-independent)."""
        llm, cfg = self.model, self.model.config
        dev = input_ids.device
        G, S = input_ids.shape
        emb = llm.model.embed(input_ids)
        image_mask = None
        if pixel_values is not None:
            emb, image_mask = self.prompt_wrap_vision(input_ids, emb, self.extract_image_feature(pixel_values))
        # a persistent workspace cache (so the captured AR-step graph stays valid across calls)
        need = S + 1 + cfg.num_image_tokens_for_gen + 8
        rows = max(3, 3 * G)
        cache = getattr(self, "_ws_cache", None)
        if cache is None or cache.max_len < need or cache.k[0].device != dev or cache.max_batch < rows:
            cache = self._ws_cache = llm.new_cache(max_len=max(need, 512), max_batch=rows)
        cache.seq_len, cache.batch = 0, G
        pos = torch.arange(S, device=dev, dtype=torch.int32).unsqueeze(0).expand(G, S)
        llm.model.forward_tokens(emb, pos, cache, key_mask=None, image_mask=image_mask)
        start = llm.model.embed(torch.full((G, 1), cfg.image_start_token, device=dev, dtype=torch.long))
        am = torch.ones((G, S + 1), dtype=torch.int32, device=dev)
        img, _, fmask = llm.generate_image(
            input_embeds=start, past_key_values=cache, attention_mask=am, uncond_attention_mask=uncond_attention_mask,
            text_uncond_attention_mask=text_uncond_attention_mask,
            latent_to_sem_func=self.vision.forward_feature_decoder, linear_proj=self.linear_proj,
            sem_to_pix_func=self.vision.forward_pixel_decoder, image_gen_temperature=image_gen_temperature,
            noises=noises)
        self.past_key_values, self.past_attention_mask = cache, fmask[0:1]
        llm.model.check_expert_parallel()
        return (img[0:1] if G == 1 else img), fmask

    @torch.no_grad()
    def generate_text(self, input_ids, pixel_values=None, max_new_tokens: int = 32, eos_token_id=None):
        """Image -> text understanding / plain text continuation (BASELINE configs[2] path): prefill `input_ids` with
        the image features scattered at the `<imagePatch>` positions (routed by `image_gate`,
        modeling_bailingmm.py:245-258), then greedy decoding — what `MingUniVisionForConditionalGeneration.generate`
        does through HF GenerationMixin with do_sample = false (config.json:30).  Stops at `eos_token_id` (default:
        config.pad_token_id = <|endoftext|>) or at the `<image>` start token (image generation is a separate call
        here).  Returns the list of new token ids."""
        llm, cfg = self.model, self.model.config
        dev = input_ids.device
        eos = cfg.pad_token_id if eos_token_id is None else eos_token_id
        S = input_ids.shape[1]
        emb = llm.model.embed(input_ids)
        image_mask = None
        if pixel_values is not None:
            emb, image_mask = self.prompt_wrap_vision(input_ids, emb, self.extract_image_feature(pixel_values))
        need = S + max_new_tokens + 8
        cache = getattr(self, "_ws_cache", None)
        if cache is None or cache.max_len < need or cache.k[0].device != dev:
            cache = self._ws_cache = llm.new_cache(max_len=max(need, 512))
            llm._gen_ws, llm._txt_ws = {}, {}
        cache.seq_len, cache.batch = 0, 1
        pos = torch.arange(S, device=dev, dtype=torch.int32).unsqueeze(0)
        hidden = llm.model.forward_tokens(emb, pos, cache, key_mask=None, image_mask=image_mask)
        # greedy decoding, one CUDA-graph replay per token (BailingMoeForCausalLM.greedy_decode); the per-token EOS check
        # reads the token on the host, as HF generate does
        out_ids = llm.greedy_decode(hidden[:, -1], cache, max_new_tokens, stop_ids=(eos, cfg.image_start_token))
        self.past_key_values = cache
        return out_ids
