"""CPU oracle for the MingTok path: a plain functional fp32 restatement of the reference's algorithm.

TEST INFRASTRUCTURE — NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs may import this module; the product package (ming_univision_b200/) never does, and its
operators raise when the CUDA library is missing instead of falling back to anything here.

Parity pinning: the reference ships no golden vectors or assertions for this path.  The oracle is
pinned against the UNMODIFIED reference modules run in the build container on identical seeded weights and inputs:
tests/golden/make_golden.py imports /root/reference (through oracle/ref_shims.py), dumps inputs/outputs to
tests/golden/*.npz, and tests/test_oracle_golden.py checks this file against those fixtures on CPU.

Every function cites the reference lines it restates (paths relative to the reference repo).  All arithmetic is fp32
(BASELINE config 1 / regime "R1 all-fp32 limit"); `state` is a flat dict with the reference's
state_dict keys.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

LN_EPS = 1e-6  # partial(nn.LayerNorm, eps=1e-6): mingtok/vision_transformer/vision_transformer.py:97,282


def _lin(sd, prefix, x):
    return F.linear(x, sd[prefix + ".weight"], sd.get(prefix + ".bias"))


def _ln(sd, prefix, x):
    return F.layer_norm(x, (x.shape[-1],), sd[prefix + ".weight"], sd[prefix + ".bias"], LN_EPS)


def _nblocks(sd, prefix):
    n = 0
    while f"{prefix}.blocks.0.{n}.norm1.weight" in sd:
        n += 1
    return n


# ---------------------------------------------------------------------------------------------------------------
# attention (eager fp32 twins; flash-attn computes the same function)
# ---------------------------------------------------------------------------------------------------------------
def attention_full(sd, prefix, x, num_heads):
    """Attention.forward — mingtok/vision_transformer/layers/attention.py:61-74 (q pre-scaled, softmax, no mask)."""
    B, N, C = x.shape
    qkv = _lin(sd, prefix + ".qkv", x).reshape(B, N, 3, num_heads, C // num_heads).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0] * (C // num_heads) ** -0.5, qkv[1], qkv[2]
    attn = (q @ k.transpose(-2, -1)).softmax(dim=-1)
    x = (attn @ v).transpose(1, 2).reshape(B, N, C)
    return _lin(sd, prefix + ".proj", x)


def attention_causal(sd, prefix, x, num_heads, cache=None):
    """CausalAttention.forward — layers/attention.py:138-163.  `cache` is None or a dict {"k","v"} of [B,H,T,hd]
    tensors that is extended in place (DynamicCache.update semantics, vision_transformer.py:396); with a non-empty
    cache the reference is only ever called with N == 1, where the triu mask is empty (attends to every key)."""
    B, N, C = x.shape
    hd = C // num_heads
    qkv = _lin(sd, prefix + ".qkv", x).reshape(B, N, 3, num_heads, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0] * hd ** -0.5, qkv[1], qkv[2]
    if cache is not None:
        if cache.get("k") is not None:
            k = torch.cat([cache["k"], k], dim=2)
            v = torch.cat([cache["v"], v], dim=2)
        cache["k"], cache["v"] = k, v
    attn = q @ k.transpose(-2, -1)
    T = k.shape[2]
    # bottom-right aligned causal mask == triu(ones(N, N), 1) when T == N and all-visible when N == 1
    mask = torch.triu(torch.ones(N, T, dtype=torch.bool), diagonal=1 + (T - N))
    attn = attn.masked_fill(mask, float("-inf")).softmax(dim=-1)
    x = (attn @ v).transpose(1, 2).reshape(B, N, C)
    return _lin(sd, prefix + ".proj", x)


# ---------------------------------------------------------------------------------------------------------------
# FFNs and blocks
# ---------------------------------------------------------------------------------------------------------------
def swiglu_ffn(sd, prefix, x):
    """SwiGLUFFN.forward — layers/swiglu_ffn.py:30-34."""
    x1, x2 = _lin(sd, prefix + ".w12", x).chunk(2, dim=-1)
    return _lin(sd, prefix + ".w3", F.silu(x1) * x2)


def gelu_mlp(sd, prefix, x):
    """Mlp.forward — layers/mlp.py:34-39 (exact-erf GELU)."""
    return _lin(sd, prefix + ".fc2", F.gelu(_lin(sd, prefix + ".fc1", x)))


def _ffn(sd, prefix, x):
    return swiglu_ffn(sd, prefix, x) if prefix + ".w12.weight" in sd else gelu_mlp(sd, prefix, x)


def block(sd, prefix, x, num_heads):
    """Block.forward (eval branch) — layers/block.py:80-105: x += attn(LN1 x); x += ffn(LN2 x)."""
    x = x + attention_full(sd, prefix + ".attn", _ln(sd, prefix + ".norm1", x), num_heads)
    return x + _ffn(sd, prefix + ".mlp", _ln(sd, prefix + ".norm2", x))


def causal_block(sd, prefix, x, num_heads, cache=None):
    """CausalBlock.forward — layers/block.py:301-327."""
    x = attention_causal(sd, prefix + ".attn", _ln(sd, prefix + ".norm1", x), num_heads, cache) + x
    return _ffn(sd, prefix + ".mlp", _ln(sd, prefix + ".norm2", x)) + x


# ---------------------------------------------------------------------------------------------------------------
# low-level encoder
# ---------------------------------------------------------------------------------------------------------------
def interpolate_pos_encoding(pos_embed, npatch, w, h, patch_size, offset=0.1):
    """VisionTransformerEncoder.interpolate_pos_encoding — vision_transformer.py:183-215 (cls position is LAST)."""
    N = pos_embed.shape[1] - 1
    if npatch == N and w == h:
        return pos_embed
    pos_embed = pos_embed.float()
    patch_pos, class_pos = pos_embed[:, :-1], pos_embed[:, -1]
    dim = pos_embed.shape[-1]
    w0, h0 = w // patch_size, h // patch_size
    M = int(math.sqrt(N))
    assert N == M * M
    sx, sy = float(w0 + offset) / M, float(h0 + offset) / M
    patch_pos = F.interpolate(patch_pos.reshape(1, M, M, dim).permute(0, 3, 1, 2), mode="bicubic", antialias=False,
                              scale_factor=(sx, sy))
    assert (w0, h0) == tuple(patch_pos.shape[-2:])
    patch_pos = patch_pos.permute(0, 2, 3, 1).reshape(1, -1, dim)
    return torch.cat((patch_pos, class_pos.unsqueeze(0)), dim=1)


def encoder_forward(sd, x, cfg, prefix="low_level_encoder"):
    """VisionTransformerEncoder.forward — vision_transformer.py:218-233 (+ PatchEmbed layers/patch_embed.py:69-82,
    forward_out_layer :173-178).  x: [B,3,H,W] -> latent [B, n+1, out_dim] (un-normalised)."""
    B, _, w, h = x.shape
    P = cfg["patch_size"]
    heads = cfg["embed_dim"] // 64  # vision_transformer.py:661
    t = F.conv2d(x, sd[prefix + ".patch_embed.proj.weight"], sd[prefix + ".patch_embed.proj.bias"], stride=P)
    t = t.flatten(2).transpose(1, 2)
    t = torch.cat((t, sd[prefix + ".cls_token"].expand(B, -1, -1)), dim=1)  # cls appended at the END (:221)
    t = t + interpolate_pos_encoding(sd[prefix + ".pos_embed"], t.shape[1] - 1, w, h, P)
    for i in range(_nblocks(sd, prefix)):
        t = block(sd, f"{prefix}.blocks.0.{i}", t, heads)
    out_dim = cfg["out_dim"]
    shortcut = t.reshape(B, t.shape[1], out_dim, -1).mean(-1)  # "b n (c h) -> b n c h", mean over h (:174)
    y = _lin(sd, prefix + ".out_proj", F.gelu(_ln(sd, prefix + ".out_norm", t)))
    return shortcut + y


# ---------------------------------------------------------------------------------------------------------------
# semantic decoder (causal) and pixel decoder
# ---------------------------------------------------------------------------------------------------------------
def decoder_in_projection(sd, x, embed_dim, prefix="semantic_decoder"):
    """TransformerDecoder.forward_in_projection_layer — vision_transformer.py:373-380."""
    rep = embed_dim // x.shape[-1]
    shortcut = x.unsqueeze(-1).repeat(1, 1, 1, rep).flatten(2)  # "b n c h -> b n (c h)"
    return _lin(sd, prefix + ".in_proj", x) + shortcut


def semantic_decoder_forward(sd, latent, cfg, caches=None, prefix="semantic_decoder"):
    """TransformerDecoder.forward_features for the causal decoder — vision_transformer.py:382-451.
    latent: [B, N, in_dim] (un-normalised).  Without caches: full causal pass, returns x_norm[:, :-1] when N > 1
    (trailing cls dropped, :431-439).  With `caches` (list of per-layer dicts) the call is incremental."""
    E = cfg["embed_dim"]
    heads = E // 64  # vision_transformer.py:618
    x = decoder_in_projection(sd, latent, E, prefix)
    N = x.shape[1]
    for i in range(_nblocks(sd, prefix)):
        x = causal_block(sd, f"{prefix}.blocks.0.{i}", x, heads, None if caches is None else caches[i])
    x = _ln(sd, prefix + ".norm", x)
    return x[:, :-1] if N > 1 else x


def new_decoder_caches(sd, prefix="semantic_decoder"):
    return [dict(k=None, v=None) for _ in range(_nblocks(sd, prefix))]


def pixel_decoder_forward(sd, feats, cfg_sem, cfg_pix):
    """MingTok.forward_pixel_decoder — mingtok/modeling_mingtok.py:179-196: sem_to_pix + pixel-shuffle rearrange,
    24 full-attention blocks without positional embedding (vision_transformer.py:572-597), LN, head, unpatchify
    (:515-527), clamp(-1, 1)."""
    B, n, _ = feats.shape
    E = cfg_pix["embed_dim"]
    f = cfg_sem["patch_size"] // cfg_pix["patch_size"]
    g = int(math.sqrt(n))
    x = _lin(sd, "sem_to_pix", feats)
    # "b (h w) (x y c) -> b (h x w y) c"
    x = x.reshape(B, g, g, f, f, E).permute(0, 1, 3, 2, 4, 5).reshape(B, g * f * g * f, E)
    heads = E // 64  # vision_transformer.py:581
    for i in range(_nblocks(sd, "pixel_decoder")):
        x = block(sd, f"pixel_decoder.blocks.0.{i}", x, heads)
    x = _lin(sd, "pixel_decoder.head", _ln(sd, "pixel_decoder.norm", x))
    p = cfg_pix["patch_size"]
    hh = int(math.sqrt(x.shape[1]))
    img = torch.einsum("nhwpqc->nchpwq", x.reshape(B, hh, hh, p, p, 3)).reshape(B, 3, hh * p, hh * p)
    return img.clamp(-1, 1)


# ---------------------------------------------------------------------------------------------------------------
# MingTok container
# ---------------------------------------------------------------------------------------------------------------
def mingtok_forward(sd, x, config):
    """MingTok.forward — modeling_mingtok.py:156-163."""
    latent = encoder_forward(sd, x, config["low_level_encoder"])
    feats = semantic_decoder_forward(sd, latent, config["semantic_decoder"])
    return {"x_norm_patchtokens": feats, "latent": (latent - config["mean"]) / config["scaling_factor"]}


def mingtok_forward_enc_dec(sd, x, config):
    """MingTok.forward_enc_dec — modeling_mingtok.py:150-153."""
    feats = mingtok_forward(sd, x, config)["x_norm_patchtokens"]
    return pixel_decoder_forward(sd, feats, config["semantic_decoder"], config["pixel_decoder"])


def mingtok_forward_feature_decoder(sd, latent_norm, config, caches):
    """MingTok.forward_feature_decoder — modeling_mingtok.py:165-174 (de-normalise, one cached causal step)."""
    hidden = latent_norm * config["scaling_factor"] + config["mean"]
    return semantic_decoder_forward(sd, hidden, config["semantic_decoder"], caches)
