"""CPU oracle for the Bailing-MoE AR path: fp32 functional restatement of mingunivision/modeling_bailing_moe.py
(BailingMoeRMSNorm :122-136, BailingMoeRotaryEmbeddingLegacy :213-237 + apply_rotary_pos_emb :428-461,
BailingMoeAttention eager :743-829, BailingMoeGate :505-520, BailingMoeSparseMoeBlock :556-639, BailingMoeMLP :483-484,
BailingMoeDecoderLayer :1165-1239, BailingMoeModel.forward :1391-1540, vis_head :1571-1574,
forward_for_image_generation_inner :1622-1673, generate_image :1844-1965).

TEST INFRASTRUCTURE — NOT PRODUCT CODE (rules in oracle/mingtok_oracle.py).  Pinned against the unmodified reference
by tests/golden/make_golden_llm.py -> tests/golden/llm_tiny.npz.  `sd` uses the reference's state_dict keys of
BailingMoeForCausalLM (model.layers.{l}.attention.query_key_value.weight, ...); `cfg` is a dict with the
BailingMoeConfig field names.  RoPE is the 1-D legacy variant (rope_scaling=None); the 3-D M-RoPE
variant is restated separately (mrope_tables / apply_mrope) for the config-gated kernel.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from . import rf_oracle


def rmsnorm(x, w, eps):
    """BailingMoeRMSNorm.forward — :131-136."""
    x32 = x.float()
    var = x32.pow(2).mean(-1, keepdim=True)
    return w * (x32 * torch.rsqrt(var + eps))


def rope_tables(head_dim, base, seq_len):
    """BailingMoeRotaryEmbeddingLegacy.forward — :213-237 (fp32 tables, emb = cat(freqs, freqs))."""
    inv_freq = 1.0 / (base ** (torch.arange(0, head_dim, 2).float() / head_dim))
    t = torch.arange(seq_len, dtype=torch.float32)
    freqs = torch.outer(t, inv_freq)
    emb = torch.cat((freqs, freqs), dim=-1)
    return emb.cos(), emb.sin()


def rotate_half(x):
    x1, x2 = x[..., : x.shape[-1] // 2], x[..., x.shape[-1] // 2:]
    return torch.cat((-x2, x1), dim=-1)


def mrope_tables(head_dim, base, position_ids3):
    """BailingMoe3DRotaryEmbedding.forward — :413-425 (the config-gated `rope_scaling.type == "3D"` variant, SURVEY.md
    §0.4): position_ids3 [3, B, S] (temporal, height, width) -> fp32 cos / sin [3, B, S, head_dim], never cast down."""
    inv_freq = 1.0 / (base ** (torch.arange(0, head_dim, 2).float() / head_dim))
    freqs = position_ids3[..., None].float() * inv_freq  # the reference's [dim/2, 1] @ [1, S] matmul: one product each
    emb = torch.cat((freqs, freqs), dim=-1)
    return emb.cos(), emb.sin()


def apply_mrope(q, k, cos, sin, mrope_section=(16, 24, 24)):
    """apply_multimodal_rotary_pos_emb — :463-469: the head dimension is cut into sections [16, 24, 24, 16, 24, 24] and
    section i takes its angles from position component i % 3.  q [B, H, S, hd], k [B, Hkv, S, hd] (bf16 or fp32);
    cos / sin fp32 [3, B, S, hd].  The products promote to fp32 and the results STAY fp32 (the attention casts them to
    the compute dtype afterwards, :946-975), i.e. one rounding at the very end."""
    sec = list(mrope_section) * 2
    cos = torch.cat([m[i % 3] for i, m in enumerate(cos.split(sec, dim=-1))], dim=-1).unsqueeze(1)
    sin = torch.cat([m[i % 3] for i, m in enumerate(sin.split(sec, dim=-1))], dim=-1).unsqueeze(1)
    return (q * cos) + (rotate_half(q) * sin), (k * cos) + (rotate_half(k) * sin)


def causal_4d_mask(attention_mask, bsz, q_len, past_len, dtype=torch.float32):
    """_prepare_4d_causal_attention_mask (transformers.modeling_attn_mask_utils) as used at :1463-1466: additive mask
    [B,1,q_len,past+q_len], bottom-right aligned causal triangle merged with the 2-D padding mask (0 = masked)."""
    kv = past_len + q_len
    minv = torch.finfo(dtype).min
    m = torch.zeros((bsz, 1, q_len, kv), dtype=dtype)
    if q_len > 1:
        causal = torch.triu(torch.ones(q_len, kv, dtype=torch.bool), diagonal=1 + past_len)
        m = m.masked_fill(causal[None, None], minv)
    if attention_mask is not None:
        pad = (attention_mask[:, None, None, :kv] == 0)
        m = m.masked_fill(pad, minv)
    return m


def attention(sd, prefix, cfg, h, mask4d, position_ids, cache):
    """BailingMoeAttention.forward (eager) — :743-829.  cache: dict(k, v) of [B, Hkv, T, hd] or None entries."""
    B, S, _ = h.shape
    H, Hkv, hd = cfg["num_attention_heads"], cfg["num_key_value_heads"], cfg["head_dim"]
    qkv = F.linear(h, sd[prefix + ".query_key_value.weight"], sd.get(prefix + ".query_key_value.bias"))
    qkv = qkv.view(B, S, H + 2 * Hkv, hd)
    q, k, v = qkv.split([H, Hkv, Hkv], dim=-2)
    q, k, v = q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2)
    past = 0 if cache.get("k") is None else cache["k"].shape[2]
    cos, sin = rope_tables(hd, cfg["rope_theta"], max(int(position_ids.max()) + 1, past + S))
    cos, sin = cos[position_ids].unsqueeze(1), sin[position_ids].unsqueeze(1)
    q = q * cos + rotate_half(q) * sin
    k = k * cos + rotate_half(k) * sin
    if cache.get("k") is not None:
        k = torch.cat([cache["k"], k], dim=2)
        v = torch.cat([cache["v"], v], dim=2)
    cache["k"], cache["v"] = k, v
    rep = H // Hkv
    kk = k[:, :, None].expand(B, Hkv, rep, k.shape[2], hd).reshape(B, H, k.shape[2], hd)
    vv = v[:, :, None].expand(B, Hkv, rep, v.shape[2], hd).reshape(B, H, v.shape[2], hd)
    w = torch.matmul(q / math.sqrt(hd), kk.transpose(2, 3))
    if mask4d is not None:
        w = w + mask4d
    w = F.softmax(w, dim=-1, dtype=torch.float32)
    o = torch.matmul(w, vv).transpose(1, 2).reshape(B, S, H * hd)
    return F.linear(o, sd[prefix + ".dense.weight"], sd.get(prefix + ".dense.bias"))


def mlp(sd, prefix, x):
    """BailingMoeMLP.forward — :483-484."""
    return F.linear(F.silu(F.linear(x, sd[prefix + ".gate_proj.weight"])) * F.linear(x, sd[prefix + ".up_proj.weight"]),
                    sd[prefix + ".down_proj.weight"])


def gate(sd, prefix, cfg, x2d):
    """BailingMoeGate.forward — :505-520."""
    logits = F.linear(x2d, sd[prefix + ".weight"])
    if cfg.get("router_logits_bf16"):  # emulate the GPU regime: F.linear under bf16 autocast returns bf16 logits (:509)
        logits = logits.to(torch.bfloat16).float()
    scores = logits.softmax(dim=-1, dtype=torch.float32)
    w, idx = torch.topk(scores, k=cfg["num_experts_per_tok"], dim=-1)
    if cfg["num_experts_per_tok"] > 1 and cfg.get("norm_topk_prob", True):
        w = w / w.sum(dim=-1, keepdim=True)
    return idx, w, logits


def moe_block(sd, prefix, cfg, h, image_mask=None):
    """BailingMoeSparseMoeBlock.forward + moe_infer — :556-639 (text gate; image gate only where image_mask is set)."""
    B, S, D = h.shape
    x = h.reshape(-1, D)
    idx, w, _ = gate(sd, prefix + ".gate", cfg, x)
    if cfg.get("multi_gate", False) and image_mask is not None:
        iidx, iw, _ = gate(sd, prefix + ".image_gate", cfg, x)
        im = image_mask.reshape(-1, 1)
        idx = idx * ~im + iidx * im
        w = w * ~im + iw * im
    y = moe_infer(sd, prefix, cfg, x, idx, w)
    if cfg.get("num_shared_experts"):
        y = y + mlp(sd, prefix + ".shared_experts", x)
    return y.view(B, S, D), idx


def moe_infer(sd, prefix, cfg, x, topk_ids, topk_weight):
    """BailingMoeSparseMoeBlock.moe_infer — :608-639, step for step: count the tokens of every expert, argsort the
    (token, slot) pairs by expert, run each hit expert ONCE on its contiguous group of rows, scatter the rows back to pair
    order, weight them in the weights' dtype (fp32), sum over the slots, cast back."""
    n_exp = cfg["num_experts"]
    cnts = topk_ids.new_zeros((topk_ids.shape[0], n_exp))
    cnts.scatter_(1, topk_ids, 1)
    tokens_per_expert = cnts.sum(dim=0).tolist()
    idxs = topk_ids.reshape(-1).argsort()
    sorted_tokens = x[idxs // topk_ids.shape[1]]
    outputs, start = [], 0
    for e, n in enumerate(tokens_per_expert):
        if n == 0:
            continue
        outputs.append(mlp(sd, f"{prefix}.experts.{e}", sorted_tokens[start:start + n]))
        start += n
    outs = torch.cat(outputs, dim=0) if outputs else sorted_tokens.new_empty(0)
    new_x = torch.empty_like(outs)
    new_x[idxs] = outs
    return (new_x.view(*topk_ids.shape, -1).type(topk_weight.dtype).mul_(topk_weight.unsqueeze(dim=-1)).sum(dim=1)
            .type(new_x.dtype))


def decoder_layer(sd, l, cfg, h, mask4d, position_ids, cache, image_mask=None):
    """BailingMoeDecoderLayer.forward — :1165-1239."""
    p = f"model.layers.{l}"
    eps = cfg["rms_norm_eps"]
    h = h + attention(sd, p + ".attention", cfg, rmsnorm(h, sd[p + ".input_layernorm.weight"], eps), mask4d,
                      position_ids, cache)
    y, _ = moe_block(sd, p + ".mlp", cfg, rmsnorm(h, sd[p + ".post_attention_layernorm.weight"], eps), image_mask)
    return h + y


def model_forward(sd, cfg, inputs_embeds, attention_mask, position_ids, caches, image_mask=None):
    """BailingMoeModel.forward — :1391-1540 (eager: 4-D mask built for every call).  caches: list of per-layer dicts,
    extended in place.  Returns the final-norm hidden states [B, S, D]."""
    B, S, _ = inputs_embeds.shape
    past = 0 if caches[0].get("k") is None else caches[0]["k"].shape[2]
    if position_ids is None:
        position_ids = torch.arange(past, past + S).unsqueeze(0)
    mask4d = causal_4d_mask(attention_mask, B, S, past)
    h = inputs_embeds
    for l in range(cfg["num_hidden_layers"]):
        h = decoder_layer(sd, l, cfg, h, mask4d, position_ids, caches[l], image_mask)
    return rmsnorm(h, sd["model.norm.weight"], cfg["rms_norm_eps"])


def new_caches(cfg):
    return [dict(k=None, v=None) for _ in range(cfg["num_hidden_layers"])]


def vis_head(sd, h):
    """vis_head = Linear + LayerNorm(eps 1e-6) — :1571-1574."""
    z = F.linear(h, sd["vis_head.0.weight"], sd["vis_head.0.bias"])
    return F.layer_norm(z, (z.shape[-1],), sd["vis_head.1.weight"], sd["vis_head.1.bias"], 1e-6)


def lm_logits(sd, h):
    """compute_logit — :1604-1620 (norm_head False)."""
    return F.linear(h, sd["lm_head.weight"]).float()


def linear_proj(sd, feats):
    """MingUniVisionForConditionalGeneration.linear_proj: Linear, GELU, Linear — modeling_bailingmm.py:111-115."""
    return F.linear(F.gelu(F.linear(feats, sd["linear_proj.0.weight"], sd["linear_proj.0.bias"])),
                    sd["linear_proj.2.weight"], sd["linear_proj.2.bias"])


def image_gen_step(sd, cfg, rf_sd, rf_steps, inputs_embeds, attention_mask, position_ids, caches, noise, temperature):
    """forward_for_image_generation_inner — :1622-1673: one LLM step on the given embeddings, z = vis_head(last hidden),
    latent = diffloss.sample(z, T, text_cfg=3.0, image_cfg=1.1) (the CFG scales are hard-wired by the kwargs bug,
    SURVEY.md §0.6)."""
    h = model_forward(sd, cfg, inputs_embeds, attention_mask, position_ids, caches)
    z = vis_head(sd, h[:, -1:, :]).reshape(h.shape[0], -1)
    lat = rf_oracle.sample(rf_sd, z, noise, rf_steps, temperature, 3.0, 1.1)
    return lat.unsqueeze(1), z


def generate_image(sd, cfg, rf_sd, rf_steps, input_embeds, caches, attention_mask, uncond_attention_mask,
                   text_uncond_attention_mask, latent_to_sem, lin_proj, noises, temperature=1.0, num_tokens=None):
    """generate_image — :1844-1965.  `caches` hold the cond-row prefill (batch 1) and are repeated to the CFG rows
    (:1891-1902) then trimmed back to row 0 (:1954-1962); `noises[i]` is the [1, C] noise of step i (replaces randn);
    latent_to_sem(latent [B,1,C], state) -> (feature [B,1,F], state); lin_proj(feature) -> [B,1,D].
    Returns (list of per-token features, latents, final attention mask)."""
    assert attention_mask.shape[0] == 1
    if uncond_attention_mask is not None:
        n_c, n_u = attention_mask.shape[1], uncond_attention_mask.shape[1]
        if n_u < n_c:
            uncond_attention_mask = torch.cat((uncond_attention_mask, attention_mask[:, n_u:]), dim=1)
        attention_mask = torch.cat((attention_mask, uncond_attention_mask), dim=0)
    if text_uncond_attention_mask is not None and text_uncond_attention_mask.sum() > 0:
        n_c, n_u = attention_mask.shape[1], text_uncond_attention_mask.shape[1]
        if n_u < n_c:
            text_uncond_attention_mask = torch.cat((text_uncond_attention_mask, attention_mask[0:1, n_u:]), dim=1)
        if (text_uncond_attention_mask == uncond_attention_mask).sum() != uncond_attention_mask.numel():
            attention_mask = torch.cat((attention_mask, text_uncond_attention_mask), dim=0)
    B = attention_mask.shape[0]
    if B > 1:
        input_embeds = input_embeds.repeat((B, 1, 1))
        for c in caches:
            c["k"], c["v"] = c["k"].repeat((B, 1, 1, 1)), c["v"].repeat((B, 1, 1, 1))
    n_tok = cfg.get("num_image_tokens_for_gen", 256) if num_tokens is None else num_tokens
    feats, lats, sem_state = [], [], None
    for i in range(n_tok + 1):
        pos = (attention_mask.long().cumsum(-1) - 1)[:, -1:]
        lat, _ = image_gen_step(sd, cfg, rf_sd, rf_steps, input_embeds, attention_mask, pos, caches, noises[i],
                                temperature)
        if i < n_tok:
            lats.append(lat)
            feat, sem_state = latent_to_sem(lat, sem_state)
            feats.append(feat)
            input_embeds = lin_proj(feat)
            attention_mask = torch.cat((attention_mask, torch.ones(B, 1, dtype=attention_mask.dtype)), dim=-1)
    for c in caches:
        c["k"], c["v"] = c["k"][0:1], c["v"][0:1]
    return feats, lats, attention_mask
