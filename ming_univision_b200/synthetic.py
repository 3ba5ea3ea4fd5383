"""Seeded synthetic weights and inputs with the reference's state_dict keys and true shapes.

No pretrained checkpoint, dataset or network exists in the build / benchmark environment, so
parity tests, smoke() and bench.py all use these factories.  Every tensor is drawn from its own CPU generator seeded
by (seed, key name), so any subset of the state_dict can be regenerated bit-identically on any machine.

Distributions are chosen so activations stay O(1) through all 60 transformer blocks (which keeps the parity tests
sensitive): Linear / conv weights ~ N(0, 1/fan_in), biases ~ N(0, 0.1^2), LayerNorm gamma ~ 1 + N(0, 0.1^2),
beta ~ N(0, 0.1^2), cls / pos-embed ~ N(0, 0.5^2); the pixel head is scaled by 0.3 so the image is not saturated by
the final clamp(-1, 1).
"""
from __future__ import annotations

import hashlib
import json
import math
import os

import torch

MINGTOK_CONFIG = {  # mingtok/config/config_mingtok.json:3-26 of the reference
    "low_level_encoder": {"img_size": 512, "patch_size": 32, "depth": 12, "embed_dim": 768,
                          "ffn_layer": "swiglufused", "out_dim": 32},
    "semantic_decoder": {"in_dim": 32, "patch_size": 32, "embed_dim": 1024, "decoder_depth": 24,
                         "ffn_layer": "swiglufused"},
    "pixel_decoder": {"patch_size": 16, "decoder_depth": 24, "norm_pix_loss": True, "embed_dim": 1024,
                      "loss_type": "L1-plain"},
    "scaling_factor": 8.09449291,
    "mean": 1.46817409,
}

MINGTOK_TINY_CONFIG = {  # same topology, small widths: used for committed golden fixtures and fast CPU tests
    "low_level_encoder": {"img_size": 128, "patch_size": 32, "depth": 2, "embed_dim": 128,
                          "ffn_layer": "swiglufused", "out_dim": 32},
    "semantic_decoder": {"in_dim": 32, "patch_size": 32, "embed_dim": 128, "decoder_depth": 2,
                         "ffn_layer": "swiglufused"},
    "pixel_decoder": {"patch_size": 16, "decoder_depth": 2, "norm_pix_loss": True, "embed_dim": 128,
                      "loss_type": "L1-plain"},
    "scaling_factor": 8.09449291,
    "mean": 1.46817409,
}


def swiglu_hidden(dim: int, mlp_ratio: float = 4.0) -> int:
    """SwiGLUFFNFused hidden width — mingtok/vision_transformer/layers/swiglu_ffn.py:66."""
    return (int(int(dim * mlp_ratio) * 2 / 3) + 7) // 8 * 8


def _gen(seed: int, key: str) -> torch.Generator:
    h = hashlib.sha256(f"{seed}:{key}".encode()).digest()
    return torch.Generator(device="cpu").manual_seed(int.from_bytes(h[:8], "little") % (2 ** 63))


def _normal(seed, key, shape, std, mean=0.0):
    return torch.randn(shape, generator=_gen(seed, key), dtype=torch.float32) * std + mean


def mingtok_param_shapes(config: dict) -> dict[str, tuple]:
    """Reference state_dict schema (vision_transformer.py:97-171, 282-368, modeling_mingtok.py:117)."""
    enc, sem, pix = config["low_level_encoder"], config["semantic_decoder"], config["pixel_decoder"]
    shapes: dict[str, tuple] = {}

    def linear(name, out_f, in_f):
        shapes[name + ".weight"] = (out_f, in_f)
        shapes[name + ".bias"] = (out_f,)

    def norm(name, d):
        shapes[name + ".weight"] = (d,)
        shapes[name + ".bias"] = (d,)

    def blocks(prefix, depth, d, ffn):
        for i in range(depth):
            p = f"{prefix}.blocks.0.{i}"
            norm(p + ".norm1", d)
            linear(p + ".attn.qkv", 3 * d, d)
            linear(p + ".attn.proj", d, d)
            norm(p + ".norm2", d)
            if ffn == "swiglu":
                hdn = swiglu_hidden(d)
                linear(p + ".mlp.w12", 2 * hdn, d)
                linear(p + ".mlp.w3", d, hdn)
            else:
                linear(p + ".mlp.fc1", 4 * d, d)
                linear(p + ".mlp.fc2", d, 4 * d)

    E, P = enc["embed_dim"], enc["patch_size"]
    npatch = (enc["img_size"] // P) ** 2
    shapes["low_level_encoder.cls_token"] = (1, 1, E)
    shapes["low_level_encoder.pos_embed"] = (1, npatch + 1, E)
    shapes["low_level_encoder.patch_embed.proj.weight"] = (E, 3, P, P)
    shapes["low_level_encoder.patch_embed.proj.bias"] = (E,)
    blocks("low_level_encoder", enc["depth"], E, "swiglu")
    norm("low_level_encoder.out_norm", E)
    linear("low_level_encoder.out_proj", enc["out_dim"], E)
    S = sem["embed_dim"]
    linear("semantic_decoder.in_proj", S, sem["in_dim"])
    blocks("semantic_decoder", sem["decoder_depth"], S, "swiglu")
    norm("semantic_decoder.norm", S)
    D = pix["embed_dim"]
    blocks("pixel_decoder", pix["decoder_depth"], D, "gelu")
    norm("pixel_decoder.norm", D)
    linear("pixel_decoder.head", pix["patch_size"] ** 2 * 3, D)
    f = sem["patch_size"] // pix["patch_size"]
    linear("sem_to_pix", D * f * f, S)
    return shapes


def mingtok_state_dict(config: dict, seed: int = 0, dtype=torch.float32) -> dict[str, torch.Tensor]:
    sd = {}
    for key, shape in mingtok_param_shapes(config).items():
        if key.endswith("cls_token") or key.endswith("pos_embed"):
            t = _normal(seed, key, shape, 0.5)
        elif ".norm" in key or "out_norm" in key:
            t = _normal(seed, key, shape, 0.1, 1.0 if key.endswith(".weight") else 0.0)
        elif key.endswith(".bias"):
            t = _normal(seed, key, shape, 0.1)
        else:
            fan_in = math.prod(shape[1:])
            std = 1.0 / math.sqrt(fan_in)
            if key.startswith("pixel_decoder.head"):
                std *= 0.3
            t = _normal(seed, key, shape, std)
        sd[key] = t.to(dtype)
    return sd


def synthetic_images(batch: int, size: int, seed: int = 1234) -> torch.Tensor:
    """ImageNet-val-shaped synthetic inputs in [-1, 1]: low-pass uniform field + N(0, 0.05^2)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    low = torch.rand((batch, 3, size // 8, size // 8), generator=g) * 2 - 1
    img = torch.nn.functional.interpolate(low, size=(size, size), mode="bicubic", align_corners=False)
    img = img + torch.randn((batch, 3, size, size), generator=g) * 0.05
    return img.clamp(-1, 1)


def save_config(config: dict, path: str) -> None:
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "w") as f:
        json.dump(config, f, indent=1)


# ---------------------------------------------------------------------------------------------------------------
# rectified-flow head (mingunivision/diff_loss_rf_swiglu.py); defaults of setup_vishead_diffloss
# (modeling_bailing_moe.py:1559-1584): width 3072, depth 12, mlp_mult 4, 16 steps, latent 32 channels
# ---------------------------------------------------------------------------------------------------------------
RF_CONFIG = {"target_channels": 32, "z_channels": 3072, "width": 3072, "depth": 12, "mlp_mult": 4,
             "num_sampling_steps": 16}
RF_TINY_CONFIG = {"target_channels": 32, "z_channels": 128, "width": 128, "depth": 2, "mlp_mult": 4,
                  "num_sampling_steps": 4}


def rf_param_shapes(cfg: dict) -> dict[str, tuple]:
    """State-dict schema of RectifiedFlowLoss (keys below `diffloss.`; diff_loss_rf_swiglu.py:295-340)."""
    W, C, Z = cfg["width"], cfg["target_channels"], cfg["z_channels"]
    hidden = (int(int(W * cfg["mlp_mult"]) * 2 / 3) + 7) // 8 * 8  # SwiGLUFFNFused, :66
    shapes: dict[str, tuple] = {}

    def linear(name, out_f, in_f):
        shapes[name + ".weight"] = (out_f, in_f)
        shapes[name + ".bias"] = (out_f,)

    linear("net.time_embed.mlp.0", W, 256)
    linear("net.time_embed.mlp.2", W, W)
    linear("net.cond_embed", W, Z)
    linear("net.input_proj", W, C)
    for i in range(cfg["depth"]):
        p = f"net.res_blocks.{i}"
        shapes[p + ".in_ln.weight"] = (W,)
        shapes[p + ".in_ln.bias"] = (W,)
        linear(p + ".mlp.w12", 2 * hidden, W)
        linear(p + ".mlp.w3", W, hidden)
        linear(p + ".adaLN_modulation.1", 3 * W, W)
    linear("net.final_layer.adaLN_modulation.1", 2 * W, W)
    linear("net.final_layer.linear", C, W)
    return shapes


def rf_state_dict(cfg: dict, seed: int = 0, dtype=torch.float32) -> dict[str, torch.Tensor]:
    """Seeded weights.  The reference zero-initialises the adaLN and output layers (diff_loss_rf_swiglu.py:353-361),
    which would make v == 0 and every test vacuous, so ALL layers are randomised: Linear ~ N(0, 1/fan_in) (adaLN
    modulation scaled by 0.5, output layer by 2 so |v| ~ 1), biases ~ N(0, 0.1^2), in_ln gamma ~ 1 + N(0, 0.1^2)."""
    sd = {}
    for key, shape in rf_param_shapes(cfg).items():
        if ".in_ln." in key:
            t = _normal(seed, "rf." + key, shape, 0.1, 1.0 if key.endswith(".weight") else 0.0)
        elif key.endswith(".bias"):
            t = _normal(seed, "rf." + key, shape, 0.1)
        else:
            std = 1.0 / math.sqrt(shape[1])
            if "adaLN_modulation" in key:
                std *= 0.5
            if key.startswith("net.final_layer.linear"):
                std *= 2.0
            t = _normal(seed, "rf." + key, shape, std)
        sd[key] = t.to(dtype)
    return sd


# ---------------------------------------------------------------------------------------------------------------
# Bailing-MoE LLM (mingunivision/config.json:11-119 llm_config; rope_scaling=None -> 1-D legacy rotary)
# ---------------------------------------------------------------------------------------------------------------
LLM_CONFIG = dict(vocab_size=126464, hidden_size=2048, intermediate_size=5632, num_hidden_layers=28,
                  num_attention_heads=16, num_key_value_heads=4, head_dim=128, hidden_act="silu", use_qkv_bias=False,
                  use_bias=False, rms_norm_eps=1e-5, max_position_embeddings=32768, rope_theta=600000,
                  rope_scaling=None, num_experts=64, num_shared_experts=2, num_experts_per_tok=6, norm_topk_prob=True,
                  moe_intermediate_size=1408, first_k_dense_replace=0, multi_gate=True, num_image_tokens_for_gen=256,
                  image_patch_token=126346, image_start_token=126347, pad_token_id=126081, embedding_dropout=0.0,
                  attention_dropout=0.0, output_dropout=0.0)
VISHEAD_CONFIG = dict(diffloss_w=3072, diffloss_d=12, num_sampling_steps="16", gen_method="flow_matching_swiglu-4",
                      hidden_size=2048, vis_head_arch="linear2-norm", image_emb_dim_for_gen=32)

# the true 16B-A3B widths with 2 layers instead of 28 (SURVEY.md §8c "layer-exact, depth-reduced"): golden fixture
# tests/golden/llm_wide.npz and the CPU baseline of bench.py (per-layer time x 28, flagged as extrapolated)
LLM_WIDE_CONFIG = dict(LLM_CONFIG, num_hidden_layers=2, num_image_tokens_for_gen=4)

LLM_TINY_CONFIG = dict(vocab_size=512, hidden_size=128, intermediate_size=256, num_hidden_layers=2,
                       num_attention_heads=4, num_key_value_heads=2, head_dim=128, hidden_act="silu",
                       use_qkv_bias=False, use_bias=False, rms_norm_eps=1e-5, max_position_embeddings=4096,
                       rope_theta=600000, rope_scaling=None, num_experts=8, num_shared_experts=1,
                       num_experts_per_tok=2, norm_topk_prob=True, moe_intermediate_size=64, first_k_dense_replace=0,
                       multi_gate=True, num_image_tokens_for_gen=4, image_patch_token=500, image_start_token=501,
                       pad_token_id=0, embedding_dropout=0.0, attention_dropout=0.0, output_dropout=0.0)
VISHEAD_TINY_CONFIG = dict(diffloss_w=128, diffloss_d=2, num_sampling_steps="4", gen_method="flow_matching_swiglu-4",
                           hidden_size=128, vis_head_arch="linear2-norm", image_emb_dim_for_gen=32)


def rf_config_from_vishead(vh: dict) -> dict:
    return {"target_channels": vh["image_emb_dim_for_gen"], "z_channels": vh["diffloss_w"], "width": vh["diffloss_w"],
            "depth": vh["diffloss_d"], "mlp_mult": int(vh["gen_method"].split("-")[1]),
            "num_sampling_steps": int(vh["num_sampling_steps"])}


def llm_param_shapes(cfg: dict, vh: dict | None = None, feature_dim: int | None = None) -> dict[str, tuple]:
    """State-dict schema of BailingMoeForCausalLM (+ vis_head, diffloss when `vh`; + the wrapper's linear_proj when
    `feature_dim`) — SURVEY.md §3.5; modeling_bailing_moe.py:479-484, 505-520, 543-552, 680-686, 1359-1389, 1543-1584."""
    D, H, Hkv, hd = cfg["hidden_size"], cfg["num_attention_heads"], cfg["num_key_value_heads"], cfg["head_dim"]
    E, I = cfg["num_experts"], cfg["moe_intermediate_size"]
    shapes: dict[str, tuple] = {"model.word_embeddings.weight": (cfg["vocab_size"], D)}
    for l in range(cfg["num_hidden_layers"]):
        p = f"model.layers.{l}"
        shapes[p + ".attention.query_key_value.weight"] = ((H + 2 * Hkv) * hd, D)
        shapes[p + ".attention.dense.weight"] = (D, H * hd)
        shapes[p + ".input_layernorm.weight"] = (D,)
        shapes[p + ".post_attention_layernorm.weight"] = (D,)
        for gname in ("gate", "image_gate", "audio_gate") if cfg.get("multi_gate") else ("gate",):
            shapes[f"{p}.mlp.{gname}.weight"] = (E, D)
        for e in range(E):
            shapes[f"{p}.mlp.experts.{e}.gate_proj.weight"] = (I, D)
            shapes[f"{p}.mlp.experts.{e}.up_proj.weight"] = (I, D)
            shapes[f"{p}.mlp.experts.{e}.down_proj.weight"] = (D, I)
        if cfg.get("num_shared_experts"):
            Is = I * cfg["num_shared_experts"]
            shapes[p + ".mlp.shared_experts.gate_proj.weight"] = (Is, D)
            shapes[p + ".mlp.shared_experts.up_proj.weight"] = (Is, D)
            shapes[p + ".mlp.shared_experts.down_proj.weight"] = (D, Is)
    shapes["model.norm.weight"] = (D,)
    shapes["lm_head.weight"] = (cfg["vocab_size"], D)
    if vh is not None:
        Z = vh["diffloss_w"]
        shapes["vis_head.0.weight"] = (Z, D)
        shapes["vis_head.0.bias"] = (Z,)
        shapes["vis_head.1.weight"] = (Z,)
        shapes["vis_head.1.bias"] = (Z,)
        for k, v in rf_param_shapes(rf_config_from_vishead(vh)).items():
            shapes["diffloss." + k] = v
    if feature_dim is not None:
        shapes["linear_proj.0.weight"] = (D, feature_dim)
        shapes["linear_proj.0.bias"] = (D,)
        shapes["linear_proj.2.weight"] = (D, D)
        shapes["linear_proj.2.bias"] = (D,)
    return shapes


def llm_tensor(key: str, shape: tuple, seed: int = 0, dtype=torch.float32, device="cpu") -> torch.Tensor:
    """One seeded tensor of the LLM schema.  Linear ~ N(0, 1/fan_in); router gates ~ N(0, (3/sqrt(D))^2) so top-k
    margins are wide (no ties in bf16); norm weights ~ 1 + N(0, 0.1^2); embeddings ~ N(0, 1)."""
    if key.startswith("diffloss."):
        raise KeyError("use rf_state_dict for the diffloss.* keys")
    if "layernorm" in key or key.endswith("norm.weight") or key.startswith("vis_head.1"):
        return _normal(seed, "llm." + key, shape, 0.1, 1.0 if key.endswith(".weight") else 0.0).to(dtype)
    if key.endswith(".bias"):
        return _normal(seed, "llm." + key, shape, 0.1).to(dtype)
    if "word_embeddings" in key:
        return _normal(seed, "llm." + key, shape, 1.0).to(dtype)
    std = 1.0 / math.sqrt(shape[-1])
    if "gate.weight" in key and "experts" not in key:
        std *= 3.0
    return _normal(seed, "llm." + key, shape, std).to(dtype)


def llm_state_dict(cfg: dict, vh: dict | None = None, feature_dim: int | None = None, seed: int = 0,
                   dtype=torch.float32) -> dict[str, torch.Tensor]:
    shapes = {k: v for k, v in llm_param_shapes(cfg, vh, feature_dim).items() if not k.startswith("diffloss.")}
    if sum(math.prod(v) for v in shapes.values()) > 50_000_000:
        # every tensor has its own generator (seeded by its key), so the keys can be drawn concurrently — torch.randn
        # releases the GIL; the true-width fixtures (3 B parameters) take ~10 s instead of a minute
        from concurrent.futures import ThreadPoolExecutor

        with ThreadPoolExecutor(max_workers=min(32, os.cpu_count() or 8)) as ex:
            sd = dict(zip(shapes, ex.map(lambda kv: llm_tensor(kv[0], kv[1], seed, dtype), shapes.items())))
    else:
        sd = {key: llm_tensor(key, shape, seed, dtype) for key, shape in shapes.items()}
    if vh is not None:
        for k, v in rf_state_dict(rf_config_from_vishead(vh), seed, dtype).items():
            sd["diffloss." + k] = v
    return sd


@torch.no_grad()
def init_on_device(model, seed: int = 0) -> None:
    """In-place seeded initialisation of a MingUniVisionForConditionalGeneration on ITS device (device-side generator):
    the same distributions as mingtok_state_dict / llm_state_dict / rf_state_dict, drawn on the GPU because 16.8 B
    parameters cannot be drawn on the host in bench time.  The values differ from the CPU factories (another RNG), so
    this is for full-size timing runs; parity runs load the CPU-seeded state dicts.  Parameters left on the meta device
    (experts of other expert-parallel ranks) are skipped; every tensor is seeded by (seed, name), so a rank's experts get
    the values they would have in the unsharded model."""
    for name, p in model.named_parameters():
        if p.is_meta:
            continue
        h = hashlib.sha256(f"{seed}:dev:{name}".encode()).digest()
        g = torch.Generator(device=p.device).manual_seed(int.from_bytes(h[:8], "little") % (2 ** 63))
        key = name.split(".", 1)[1] if name.split(".", 1)[0] in ("vision", "model") else name

        def normal(std, mean=0.0):
            p.copy_(torch.randn(p.shape, generator=g, device=p.device, dtype=torch.float32) * std + mean)

        if name.startswith("vision."):
            if key.endswith("cls_token") or key.endswith("pos_embed"):
                normal(0.5)
            elif ".norm" in key or "out_norm" in key:
                normal(0.1, 1.0 if key.endswith(".weight") else 0.0)
            elif key.endswith(".bias"):
                normal(0.1)
            else:
                normal((0.3 if key.startswith("pixel_decoder.head") else 1.0) / math.sqrt(math.prod(p.shape[1:])))
        elif name.startswith("model.diffloss."):
            if ".in_ln." in key:
                normal(0.1, 1.0 if key.endswith(".weight") else 0.0)
            elif key.endswith(".bias"):
                normal(0.1)
            else:
                std = 1.0 / math.sqrt(p.shape[1])
                if "adaLN_modulation" in key:
                    std *= 0.5
                if "final_layer.linear" in key:
                    std *= 2.0
                normal(std)
        else:  # Bailing-MoE, vis_head, linear_proj (llm_tensor's rules)
            if "layernorm" in key or key.endswith("norm.weight") or key.startswith("vis_head.1"):
                normal(0.1, 1.0 if key.endswith(".weight") else 0.0)
            elif key.endswith(".bias"):
                normal(0.1)
            elif "word_embeddings" in key:
                normal(1.0)
            else:
                std = 1.0 / math.sqrt(p.shape[-1])
                if "gate.weight" in key and "experts" not in key:
                    std *= 3.0
                normal(std)
