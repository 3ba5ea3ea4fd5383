"""B200-native MingTok: same class / method / state_dict surface as the reference's
``mingtok/modeling_mingtok.py`` (MingTok :97-206, MingTokConfig :56-89) and
``mingtok/vision_transformer/vision_transformer.py`` (VisionTransformerEncoder :50-233, TransformerDecoder :235-570),
with every operator executed by the sm_100a kernels in libmingb200.so (ming_univision_b200/ops.py).

The nn.Module tree below exists only to own parameters under the reference's state_dict keys
(``low_level_encoder.blocks.0.{i}.attn.qkv.weight`` ... — SURVEY.md §3.5); no torch.nn forward is ever called.
Weights are re-laid-out once per device (bf16, SwiGLU gate/up rows interleaved in 128-row blocks, hidden width padded
to a multiple of 128 so UMMA tiles stay aligned) by ``_pack()``.

Numerics: bf16 operands, fp32 accumulation and fp32 LayerNorm / softmax statistics, bf16 residual stream — the
reference's "R2" regime (bf16 parameters under bf16 autocast).  Inference only.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn as nn
from transformers import PretrainedConfig, PreTrainedModel

from .. import _packs, ops

BF16 = torch.bfloat16


class MingTokConfig(PretrainedConfig):
    """Field-compatible with the reference's MingTokConfig (modeling_mingtok.py:56-89)."""
    model_type = "mingtok"

    def __init__(self, low_level_encoder=None, semantic_decoder=None, pixel_decoder=None, pretrained_checkpoint=None,
                 model_dtype="bf16", scaling_factor=1.0, mean=0.0, **kwargs):
        super().__init__(**kwargs)
        self.low_level_encoder = low_level_encoder or {}
        self.semantic_decoder = semantic_decoder or {}
        self.pixel_decoder = pixel_decoder or {}
        self.pretrained_checkpoint = pretrained_checkpoint
        self.model_dtype = model_dtype
        self.scaling_factor = scaling_factor
        self.mean = mean


def _swiglu_hidden(dim: int, mlp_ratio: float = 4.0) -> int:
    return (int(int(dim * mlp_ratio) * 2 / 3) + 7) // 8 * 8  # layers/swiglu_ffn.py:66


# ---------------------------------------------------------------------------------------------------------------
# parameter containers (names == reference attribute names)
# ---------------------------------------------------------------------------------------------------------------
class _Attention(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.qkv = nn.Linear(dim, dim * 3, bias=True)
        self.proj = nn.Linear(dim, dim, bias=True)


class _SwiGLUFFN(nn.Module):
    def __init__(self, dim):
        super().__init__()
        h = _swiglu_hidden(dim)
        self.w12 = nn.Linear(dim, 2 * h, bias=True)
        self.w3 = nn.Linear(h, dim, bias=True)


class _Mlp(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.fc1 = nn.Linear(dim, 4 * dim, bias=True)
        self.fc2 = nn.Linear(4 * dim, dim, bias=True)


class _Block(nn.Module):
    def __init__(self, dim, ffn):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.attn = _Attention(dim)
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)
        self.mlp = _SwiGLUFFN(dim) if ffn == "swiglu" else _Mlp(dim)


class _PatchEmbed(nn.Module):
    def __init__(self, patch_size, embed_dim):
        super().__init__()
        self.proj = nn.Conv2d(3, embed_dim, kernel_size=patch_size, stride=patch_size)


def _block_chunks(depth, dim, ffn):
    # BlockChunk nesting of the reference (vision_transformer.py:152-159, block_chunks=1) gives keys "blocks.0.{i}"
    return nn.ModuleList([nn.ModuleList([_Block(dim, ffn) for _ in range(depth)])])


class VisionTransformerEncoder(nn.Module):
    """Parameter layout of the reference's VisionTransformerEncoder (vision_transformer.py:97-171)."""

    def __init__(self, img_size, patch_size, embed_dim, depth, out_dim):
        super().__init__()
        self.embed_dim = self.num_features = embed_dim
        self.patch_size = patch_size
        self.num_heads = embed_dim // 64
        self.out_dim = out_dim
        self.interpolate_offset = 0.1
        self.patch_embed = _PatchEmbed(patch_size, embed_dim)
        num_patches = (img_size // patch_size) ** 2
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, num_patches + 1, embed_dim))
        self.blocks = _block_chunks(depth, embed_dim, "swiglu")
        self.out_norm = nn.LayerNorm(embed_dim, eps=1e-6)
        self.out_proj = nn.Linear(embed_dim, out_dim)


class TransformerDecoder(nn.Module):
    """Parameter layout of the reference's TransformerDecoder (vision_transformer.py:282-368)."""

    def __init__(self, patch_size, embed_dim, depth, ffn, in_dim=None, require_head=False, causal=False):
        super().__init__()
        self.embed_dim = self.num_features = embed_dim
        self.patch_size = patch_size
        self.num_heads = embed_dim // 64
        self.in_dim = in_dim
        self.causal = causal
        if in_dim is not None:
            self.in_proj = nn.Linear(in_dim, embed_dim)
        self.blocks = _block_chunks(depth, embed_dim, ffn)
        self.norm = nn.LayerNorm(embed_dim, eps=1e-6)
        self.require_head = require_head
        if require_head:
            self.head = nn.Linear(embed_dim, patch_size ** 2 * 3, bias=True)


class MingTokKVCache:
    """Static KV cache of the causal semantic decoder: per layer K and V of shape [B, H, Tmax, 64] (bf16), filled at
    position `seq_len`.  Stands in for the HF DynamicCache the reference threads through
    ``forward_feature_decoder(..., past_key_values=...)`` (vision_transformer.py:396, layers/attention.py:222-229)."""

    def __init__(self, layers: int, batch: int, heads: int, max_len: int, device):
        self.k = [torch.zeros((batch, heads, max_len, 64), dtype=BF16, device=device) for _ in range(layers)]
        self.v = [torch.zeros((batch, heads, max_len, 64), dtype=BF16, device=device) for _ in range(layers)]
        self.seq_len = 0
        self.max_len = max_len
        self.batch = batch
        # device-side copy of seq_len for CUDA-graph replay of the decode step (kept in sync by the graph itself)
        self.t_dev = torch.zeros((1,), dtype=torch.int32, device=device)

    def get_seq_length(self, layer_idx: int = 0) -> int:
        return self.seq_len

    def reset(self) -> None:
        self.seq_len = 0
        self.t_dev.zero_()


# ---------------------------------------------------------------------------------------------------------------
# packed (device-resident, kernel-ready) weights
# ---------------------------------------------------------------------------------------------------------------
def _dev_bf16(t: torch.Tensor, device) -> torch.Tensor:
    return t.detach().to(device=device, dtype=BF16).contiguous()


class _PackedBlock:
    __slots__ = ("n1w", "n1b", "qkv_w", "qkv_b", "proj_w", "proj_b", "n2w", "n2b", "ffn", "w1", "b1", "w2", "b2",
                 "w12_ref", "b12_ref", "w3_ref", "qkv_f", "w1_f")

    def __init__(self, blk: _Block, device):
        d = lambda t: _dev_bf16(t, device)  # noqa: E731
        self.n1w, self.n1b = d(blk.norm1.weight), d(blk.norm1.bias)
        self.qkv_w, self.qkv_b = d(blk.attn.qkv.weight), d(blk.attn.qkv.bias)
        self.proj_w, self.proj_b = d(blk.attn.proj.weight), d(blk.attn.proj.bias)
        self.n2w, self.n2b = d(blk.norm2.weight), d(blk.norm2.bias)
        if isinstance(blk.mlp, _SwiGLUFFN):
            self.ffn = "swiglu"
            # reference layout kept for the q_len = 1 decode step (HBM-streaming kernel), packed layout for the GEMM
            self.w12_ref, self.b12_ref, self.w3_ref = d(blk.mlp.w12.weight), d(blk.mlp.w12.bias), d(blk.mlp.w3.weight)
            self.w1, self.b1, hp = ops.pack_swiglu(self.w12_ref, self.b12_ref)
            self.w2, self.b2 = ops.pad_cols(self.w3_ref, hp), d(blk.mlp.w3.bias)
        else:
            self.ffn = "gelu"
            self.w1, self.b1 = d(blk.mlp.fc1.weight), d(blk.mlp.fc1.bias)
            self.w2, self.b2 = d(blk.mlp.fc2.weight), d(blk.mlp.fc2.bias)
        # LayerNorm-folded packs for the full-sequence path: norm1 -> qkv, norm2 -> w12 / fc1 (ops.fold_layernorm)
        self.qkv_f = ops.fold_layernorm(self.qkv_w, self.qkv_b, self.n1w, self.n1b)
        if self.ffn == "swiglu":
            H = self.w12_ref.shape[0] // 2
            wf, cs, bf = ops.fold_layernorm(self.w12_ref, self.b12_ref, self.n2w, self.n2b)
            wp, _, hp = ops.pack_swiglu(wf, None)
            self.w1_f = (wp, ops.pack_swiglu_f32(cs, H, hp), ops.pack_swiglu_f32(bf, H, hp))
        else:
            self.w1_f = ops.fold_layernorm(self.w1, self.b1, self.n2w, self.n2b)


LN_EPS = 1e-6
FOLD_LAYERNORM = True  # full-sequence blocks: LayerNorms folded into the GEMMs (no LayerNorm kernels, see _run_block_folded)


def _run_block_folded(pb: _PackedBlock, x: torch.Tensor, B: int, S: int, H: int, causal: bool,
                      stats: torch.Tensor) -> torch.Tensor:
    """The same block with both LayerNorms folded into the GEMMs that consume them: `stats` [rows, 2] holds the per-row
    (sum, sum of squares) of x on entry (left there by the previous block's last GEMM or by ops.row_stats); the proj
    and w3 / fc2 epilogues refresh it while they write the residual stream.  4 GEMMs + attention per block, no
    LayerNorm kernel, no normalised copy of the activations in HBM / L2.  Returns the statistics of the new x."""
    qkv = ops.linear(x, pb.qkv_f[0], None, ln_fold=(stats, pb.qkv_f[1], pb.qkv_f[2], LN_EPS))
    a = ops.attention_hd64(qkv, B, S, H, causal)
    rows, D = x.shape[0] * x.shape[1], x.shape[-1]
    st_mid = torch.empty((rows, (D + 63) // 64, 2), dtype=torch.float32, device=x.device)
    ops.linear(a, pb.proj_w, pb.proj_b, epi=ops.EPI_RESIDUAL, residual=x, out=x, stats_out=st_mid)
    hid = ops.linear(x, pb.w1_f[0], None, epi=ops.EPI_SWIGLU if pb.ffn == "swiglu" else ops.EPI_GELU,
                     ln_fold=(st_mid, pb.w1_f[1], pb.w1_f[2], LN_EPS))
    st_out = torch.empty_like(st_mid)
    ops.linear(hid, pb.w2, pb.b2, epi=ops.EPI_RESIDUAL, residual=x, out=x, stats_out=st_out)
    return st_out


def _run_blocks(blocks, x: torch.Tensor, B: int, S: int, H: int, causal: bool) -> torch.Tensor:
    if not FOLD_LAYERNORM:
        for pb in blocks:
            _run_block(pb, x, B, S, H, causal)
        return x
    stats = ops.row_stats(x)
    for pb in blocks:
        stats = _run_block_folded(pb, x, B, S, H, causal, stats)
    return x


def _run_block(pb: _PackedBlock, x: torch.Tensor, B: int, S: int, H: int, causal: bool) -> torch.Tensor:
    """Block.forward / CausalBlock.forward (layers/block.py:80-105, 301-327): pre-LN attention + FFN, residuals fused
    into the proj / w3 / fc2 GEMM epilogues (written in place into the residual stream x)."""
    h = ops.layernorm(x, pb.n1w, pb.n1b)
    qkv = ops.linear(h, pb.qkv_w, pb.qkv_b)
    a = ops.attention_hd64(qkv, B, S, H, causal)
    ops.linear(a, pb.proj_w, pb.proj_b, epi=ops.EPI_RESIDUAL, residual=x, out=x)
    h = ops.layernorm(x, pb.n2w, pb.n2b)
    hid = ops.linear(h, pb.w1, pb.b1, epi=ops.EPI_SWIGLU if pb.ffn == "swiglu" else ops.EPI_GELU)
    ops.linear(hid, pb.w2, pb.b2, epi=ops.EPI_RESIDUAL, residual=x, out=x)
    return x


def _run_block_step(pb: _PackedBlock, x: torch.Tensor, kc: torch.Tensor, vc: torch.Tensor, t: int,
                    t_dev: Optional[torch.Tensor] = None) -> torch.Tensor:
    """One cached decode step of a CausalBlock: x is [B, D] (q_len == 1, B = CFG rows <= 8): every linear is a
    weight-streaming pass (mb_gemv_bf16), HBM-bound."""
    if x.shape[0] > 8:
        h = ops.layernorm(x, pb.n1w, pb.n1b)
        qkv = ops.linear(h, pb.qkv_w, pb.qkv_b)
        a = ops.attention_hd64_decode(qkv, kc, vc, t, t_dev)
        ops.linear(a, pb.proj_w, pb.proj_b, epi=ops.EPI_RESIDUAL, residual=x, out=x)
        h = ops.layernorm(x, pb.n2w, pb.n2b)
        hid = ops.linear(h, pb.w1, pb.b1, epi=ops.EPI_SWIGLU)
        ops.linear(hid, pb.w2, pb.b2, epi=ops.EPI_RESIDUAL, residual=x, out=x)
        return x
    h = ops.layernorm(x, pb.n1w, pb.n1b)
    qkv = ops.gemv(h, pb.qkv_w, pb.qkv_b)
    a = ops.attention_hd64_decode(qkv, kc, vc, t, t_dev)
    ops.gemv(a, pb.proj_w, pb.proj_b, epi=ops.EPI_RESIDUAL, residual=x, out=x)
    h = ops.layernorm(x, pb.n2w, pb.n2b)
    hid = ops.gemv(h, pb.w12_ref, pb.b12_ref, epi=ops.EPI_SWIGLU)
    ops.gemv(hid, pb.w3_ref, pb.b2, epi=ops.EPI_RESIDUAL, residual=x, out=x)
    return x


class _Packed:
    """All MingTok weights in kernel-ready form on one device."""

    def __init__(self, m: "MingTok", device):
        d = lambda t: _dev_bf16(t, device)  # noqa: E731
        enc, sem, pix = m.low_level_encoder, m.semantic_decoder, m.pixel_decoder
        self.device = device
        self.pe_w = d(enc.patch_embed.proj.weight.flatten(1))  # [E, 3*P*P], column order (c, py, px)
        self.pe_b = d(enc.patch_embed.proj.bias)
        self.cls = d(enc.cls_token.reshape(-1))
        self.pos_embed_param = enc.pos_embed.detach().to(device=device, dtype=BF16)
        self.pos_cache: dict[tuple[int, int], tuple[torch.Tensor, torch.Tensor]] = {}
        self.enc_blocks = [_PackedBlock(b, device) for b in enc.blocks[0]]
        self.out_nw, self.out_nb = d(enc.out_norm.weight), d(enc.out_norm.bias)
        self.out_w, self.out_b = d(enc.out_proj.weight), d(enc.out_proj.bias)
        self.in_w, self.in_b = d(sem.in_proj.weight), d(sem.in_proj.bias)
        self.sem_blocks = [_PackedBlock(b, device) for b in sem.blocks[0]]
        self.sem_nw, self.sem_nb = d(sem.norm.weight), d(sem.norm.bias)
        self.s2p_w, self.s2p_b = d(m.sem_to_pix.weight), d(m.sem_to_pix.bias)
        self.pix_blocks = [_PackedBlock(b, device) for b in pix.blocks[0]]
        self.pix_nw, self.pix_nb = d(pix.norm.weight), d(pix.norm.bias)
        self.head_w, self.head_b = d(pix.head.weight), d(pix.head.bias)

    def pos_embed(self, w: int, h: int, patch: int, offset: float):
        """interpolate_pos_encoding (vision_transformer.py:183-215), cached per resolution: returns
        (patch positions [n, E], cls position [E]) in bf16.  The bicubic resample of the 257-row table is a
        load-time constant, computed once per (w, h) with torch on the device and reused by every forward."""
        key = (w, h)
        if key not in self.pos_cache:
            pe = self.pos_embed_param
            N = pe.shape[1] - 1
            w0, h0 = w // patch, h // patch
            if w0 * h0 == N and w == h:
                patch_pos, cls_pos = pe[0, :-1], pe[0, -1]
            else:
                pe32 = pe.float()
                M = int(math.sqrt(N))
                assert N == M * M
                sx, sy = float(w0 + offset) / M, float(h0 + offset) / M
                pp = torch.nn.functional.interpolate(pe32[:, :-1].reshape(1, M, M, -1).permute(0, 3, 1, 2),
                                                     mode="bicubic", antialias=False, scale_factor=(sx, sy))
                assert (w0, h0) == tuple(pp.shape[-2:])
                patch_pos = pp.permute(0, 2, 3, 1).reshape(w0 * h0, -1).to(BF16)
                cls_pos = pe32[0, -1].to(BF16)
            self.pos_cache[key] = (patch_pos.contiguous(), cls_pos.contiguous())
        return self.pos_cache[key]


# ---------------------------------------------------------------------------------------------------------------
# MingTok
# ---------------------------------------------------------------------------------------------------------------
class MingTok(PreTrainedModel):
    config_class = MingTokConfig
    base_model_prefix = "mingtok"

    def __init__(self, config: MingTokConfig):
        super().__init__(config)
        self.config = config
        enc, sem, pix = config.low_level_encoder, config.semantic_decoder, config.pixel_decoder
        self.latent_dim = enc.get("out_dim", 32)
        self.feature_dim = sem.get("embed_dim", 1024)
        self.patch_size = enc.get("patch_size", 32)
        self.low_level_encoder = VisionTransformerEncoder(
            img_size=enc.get("img_size", 224), patch_size=enc.get("patch_size", 16),
            embed_dim=enc.get("embed_dim", 1024), depth=enc.get("depth", 24), out_dim=enc.get("out_dim", None))
        self.semantic_decoder = TransformerDecoder(
            patch_size=sem.get("patch_size", 16), embed_dim=sem.get("embed_dim", 1024),
            depth=sem.get("decoder_depth", 1), ffn="swiglu", in_dim=sem.get("in_dim", None), causal=True)
        self.pixel_decoder = TransformerDecoder(
            patch_size=pix.get("patch_size", 16), embed_dim=pix.get("embed_dim", 1024),
            depth=pix.get("decoder_depth", 1), ffn="gelu", require_head=True)
        f = self.semantic_decoder.patch_size // self.pixel_decoder.patch_size
        self.sem_to_pix = nn.Linear(self.semantic_decoder.num_features, self.pixel_decoder.num_features * f * f)
        self.scaling_factor = config.scaling_factor
        self.mean = config.mean
        self._packed: Optional[_Packed] = None
        _packs.watch(self, self._reset_packs)  # a state-dict load through ANY ancestor drops the packed copy
        for p in self.parameters():
            p.requires_grad_(False)
        self.post_init()  # modeling_mingtok.py:126 (HF bookkeeping; _init_weights is a no-op here)

    # -- HF plumbing -------------------------------------------------------------------------------------------
    def _init_weights(self, module):  # weights always come from a checkpoint / state_dict
        return

    def _reset_packs(self) -> None:
        self._packed = None

    def _apply(self, fn, *args, **kwargs):  # .to() / .cuda() / .bfloat16() invalidate the packed copy
        self._reset_packs()
        _packs.bump()
        return super()._apply(fn, *args, **kwargs)

    @property
    def device(self):
        for _, p in self.named_parameters():
            return p.device

    def _pack(self) -> _Packed:
        dev = self.device
        if dev.type != "cuda":
            raise RuntimeError("MingTok (B200-native) runs on CUDA only: move the model to a B200 device first; "
                               "there is no CPU fallback")
        if self._packed is None or self._packed.device != dev:
            self._packed = _Packed(self, dev)
        return self._packed

    # -- stages ------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def _encode(self, x: torch.Tensor) -> torch.Tensor:
        """VisionTransformerEncoder.forward (vision_transformer.py:218-233): [B,3,H,W] -> latent [B, n+1, 32]."""
        pk = self._pack()
        enc = self.low_level_encoder
        B, _, w, h = x.shape
        P, E = enc.patch_size, enc.embed_dim
        n = (w // P) * (h // P)
        patch_pos, cls_pos = pk.pos_embed(w, h, P, enc.interpolate_offset)
        rows = ops.patchify(x, P)
        t = torch.empty((B, n + 1, E), dtype=BF16, device=x.device)
        # conv-as-GEMM + bias + pos-embed; rows of image b land at b*(n+1)+p, leaving the trailing cls slot free
        ops.linear(rows, pk.pe_w, pk.pe_b, epi=ops.EPI_RESIDUAL, residual=patch_pos, res_row_mod=n, out=t,
                   out_row_group=n, out_row_pad=1)
        ops.fill_cls_row(t, pk.cls, cls_pos)
        _run_blocks(pk.enc_blocks, t, B, n + 1, enc.num_heads, False)
        shortcut = ops.group_mean(t, enc.out_dim)
        hgelu = ops.layernorm(t, pk.out_nw, pk.out_nb, act=1)
        return ops.linear(hgelu, pk.out_w, pk.out_b, epi=ops.EPI_RESIDUAL, residual=shortcut)

    @torch.no_grad()
    def _semantic_full(self, latent: torch.Tensor) -> torch.Tensor:
        """TransformerDecoder.forward_features, use_cache=False (vision_transformer.py:382-451): un-normalised latent
        [B, N, 32] -> x_norm_patchtokens [B, N-1, 1024] (trailing cls dropped) or [B, 1, 1024] when N == 1."""
        pk = self._pack()
        sem = self.semantic_decoder
        B, N, _ = latent.shape
        x = ops.inproj_repeat(latent, pk.in_w, pk.in_b)
        _run_blocks(pk.sem_blocks, x, B, N, sem.num_heads, True)
        return ops.layernorm(x, pk.sem_nw, pk.sem_nb, drop_last_token=N > 1)

    # -- reference API -----------------------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, x):
        """modeling_mingtok.py:156-163."""
        latent = self._encode(x)
        feats = self._semantic_full(latent)
        return {"x_norm_patchtokens": feats,
                "latent": ops.affine(latent, 1.0 / self.scaling_factor, -self.mean / self.scaling_factor)}

    @torch.no_grad()
    def forward_enc_dec(self, x):
        """modeling_mingtok.py:150-153."""
        return self.forward_pixel_decoder(self.forward(x)["x_norm_patchtokens"])

    @torch.no_grad()
    def forward_feature_decoder(self, hidden_states, past_key_values: Optional[MingTokKVCache] = None):
        """modeling_mingtok.py:165-174: de-normalise the latent and run ONE cached causal step (q_len == 1)."""
        pk = self._pack()
        sem = self.semantic_decoder
        B, N, _ = hidden_states.shape
        if N != 1:
            raise ValueError("forward_feature_decoder decodes one latent token per call (the reference's AR loop)")
        cache = past_key_values
        if cache is None:
            cache = MingTokKVCache(len(pk.sem_blocks), B, sem.num_heads, 264, hidden_states.device)
        if cache.batch != B:
            raise ValueError("KV cache batch size mismatch")
        t = cache.seq_len
        if t >= cache.max_len:
            raise ValueError(f"semantic-decoder KV cache is full ({cache.max_len} tokens)")
        xn = self._decode_step(hidden_states.reshape(B, -1), cache, t, None)
        cache.seq_len = t + 1
        cache.t_dev.fill_(t + 1)
        return {"x_prenorm": None, "x_norm_patchtokens": xn.view(B, 1, -1), "past_key_values": cache}

    @torch.no_grad()
    def _decode_step(self, latent_norm: torch.Tensor, cache: MingTokKVCache, t: int,
                     t_dev: Optional[torch.Tensor]) -> torch.Tensor:
        """One cached causal step on a normalised latent [B, C] (fp32 or bf16) at position t (+ *t_dev): fixed shapes,
        no host sync — capturable in a CUDA graph (the AR-step graph of BailingMoeForCausalLM.generate_image)."""
        pk = self._pack()
        lat = ops.affine(latent_norm, self.scaling_factor, self.mean)
        x = ops.inproj_repeat(lat, pk.in_w, pk.in_b)
        for i, pb in enumerate(pk.sem_blocks):
            _run_block_step(pb, x, cache.k[i], cache.v[i], t, t_dev)
        return ops.layernorm(x, pk.sem_nw, pk.sem_nb)

    def new_decode_cache(self, batch: int, max_len: int = 264) -> MingTokKVCache:
        pk = self._pack()
        return MingTokKVCache(len(pk.sem_blocks), batch, self.semantic_decoder.num_heads, max_len, pk.device)

    @torch.no_grad()
    def forward_feature_decoder_wo_cache(self, hidden_states):
        """modeling_mingtok.py:176-177 (no de-normalisation there either)."""
        if hidden_states.dtype != BF16:
            hidden_states = ops.affine(hidden_states, 1.0, 0.0)
        return {"x_norm_patchtokens": self._semantic_full(hidden_states)}

    @torch.no_grad()
    def forward_pixel_decoder(self, x, out_dtype: torch.dtype = BF16):
        """modeling_mingtok.py:179-196: sem_to_pix + pixel shuffle, 24 full-attention blocks, LN, head, unpatchify,
        clamp(-1, 1).  x: [B, n, 1024] -> [B, 3, H, W] (bf16 / fp32), or with out_dtype=torch.uint8 the finished
        [B, H, W, 3] pixels of `tensor_to_pil` (modeling_bailing_moe.py:84-90) straight from the head GEMM's rows."""
        pk = self._pack()
        pix, sem = self.pixel_decoder, self.semantic_decoder
        if x.dtype != BF16:
            x = ops.affine(x, 1.0, 0.0)  # fp32 features (LayerNorm output under autocast in the reference)
        B, n, _ = x.shape
        g = int(math.sqrt(n))
        if g * g != n:
            raise ValueError(f"pixel decoder expects a square token grid, got {n} tokens")
        f = sem.patch_size // pix.patch_size
        s = ops.linear(x, pk.s2p_w, pk.s2p_b)
        t = ops.pixel_shuffle(s, g, f, pix.embed_dim)
        S = g * f * g * f
        _run_blocks(pk.pix_blocks, t, B, S, pix.num_heads, False)
        hn = ops.layernorm(t, pk.pix_nw, pk.pix_nb)
        y = ops.linear(hn, pk.head_w, pk.head_b)
        if out_dtype == torch.uint8:  # [B, H, W, 3] pixels as tensor_to_pil would produce them (0.5 / 0.5 statistics)
            return ops.unpatchify_to_u8(y, g * f, pix.patch_size)
        return ops.unpatchify_clamp(y, g * f, pix.patch_size, out_dtype)
