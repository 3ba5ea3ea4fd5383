"""B200-native Bailing-MoE AR path: same class names, parameter tree (state_dict keys) and method signatures as the
reference's ``mingunivision/modeling_bailing_moe.py`` for the rows of SURVEY.md §8(a) a12–a20:

  BailingMoeRMSNorm :122-136 · BailingMoeMLP :471-484 · BailingMoeGate :487-520 · BailingMoeSparseMoeBlock :523-639 ·
  BailingMoeAttention :656-829 · BailingMoeDecoderLayer :1150-1239 · BailingMoeModel :1359-1540 ·
  BailingMoeForCausalLM :1543-1965 (setup_vishead_diffloss, compute_logit, forward_for_image_generation_inner,
  generate_image).

Every operator runs in libmingb200.so (ops.py); the nn.Modules only own parameters.  Token counts on this path are
tiny (CFG rows B <= 3 per step, prompts of tens of tokens), so all linears go through the HBM-streaming kernel for
<= 8 rows and through the tcgen05 GEMM otherwise; routed experts are grouped by expert on the device (no host sync —
the reference's `tokens_per_expert.cpu()` at :616 is gone).  1-D legacy RoPE (rope_scaling=None); the
3-D M-RoPE variant is config-gated (forward_tokens with position_ids [3, B, S]).
KV caches are static [Bmax, Hkv, Tmax, hd] tensors written in place.  Inference only, bf16.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn
from transformers import PretrainedConfig

from . import _lib, _packs, ops
from .diff_loss_rf_swiglu import _FUSED_ENV, FUSED_NORM, RectifiedFlowLoss

BF16 = torch.bfloat16


class BailingMoeConfig(PretrainedConfig):
    """Field-compatible with mingunivision/configuration_bailing_moe.py:6-84."""
    model_type = "bailing_moe"

    def __init__(self, vocab_size=30592, hidden_size=1024, intermediate_size=None, num_hidden_layers=24,
                 num_attention_heads=16, num_key_value_heads=0, hidden_act="silu", use_qkv_bias=False, use_bias=True,
                 rms_norm_eps=1e-05, norm_head=False, tie_word_embeddings=False, embedding_dropout=0.1,
                 attention_dropout=0.1, output_dropout=0.1, initializer_range=0.02, max_position_embeddings=16384,
                 rope_theta=10000.0, use_cache=True, use_sliding_window=False, sliding_window=4096,
                 max_window_layers=28, rope_scaling=None, pad_token_id=126081, num_experts=16, num_shared_experts=0,
                 num_experts_per_tok=2, num_image_tokens_for_gen=256, norm_topk_prob=True, moe_intermediate_size=None,
                 first_k_dense_replace=0, head_dim=None, output_router_logits=False, multi_gate=False,
                 image_patch_token=126346, image_start_token=126347, **kwargs):
        self.num_hidden_layers = num_hidden_layers
        self.vocab_size = vocab_size
        self.hidden_size = hidden_size
        self.intermediate_size = intermediate_size
        self.num_attention_heads = num_attention_heads
        self.num_key_value_heads = num_key_value_heads
        self.hidden_act = hidden_act
        self.use_qkv_bias = use_qkv_bias
        self.use_bias = use_bias
        self.norm_head = norm_head
        self.rms_norm_eps = rms_norm_eps
        self.embedding_dropout = embedding_dropout
        self.attention_dropout = attention_dropout
        self.output_dropout = output_dropout
        self.initializer_range = initializer_range
        self.max_position_embeddings = max_position_embeddings
        self.rope_theta = rope_theta
        self.use_cache = use_cache
        self.use_sliding_window = use_sliding_window
        self.sliding_window = sliding_window
        self.max_window_layers = max_window_layers
        self.head_dim = head_dim or self.hidden_size // self.num_attention_heads
        self.rope_scaling = rope_scaling
        self.num_experts = num_experts
        self.num_shared_experts = num_shared_experts
        self.num_experts_per_tok = num_experts_per_tok
        self.num_image_tokens_for_gen = num_image_tokens_for_gen
        self.norm_topk_prob = norm_topk_prob
        self.moe_intermediate_size = moe_intermediate_size
        self.first_k_dense_replace = first_k_dense_replace
        self.output_router_logits = output_router_logits
        self.multi_gate = multi_gate
        self.image_patch_token = image_patch_token
        self.image_start_token = image_start_token
        super().__init__(pad_token_id=pad_token_id, tie_word_embeddings=tie_word_embeddings, **kwargs)


# ---------------------------------------------------------------------------------------------------------------
# parameter containers
# ---------------------------------------------------------------------------------------------------------------
class BailingMoeRMSNorm(nn.Module):
    def __init__(self, hidden_size, eps=1e-6):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(hidden_size))
        self.variance_epsilon = eps


class BailingMoeMLP(nn.Module):
    def __init__(self, config, intermediate_size):
        super().__init__()
        self.gate_proj = nn.Linear(config.hidden_size, intermediate_size, bias=False)
        self.up_proj = nn.Linear(config.hidden_size, intermediate_size, bias=False)
        self.down_proj = nn.Linear(intermediate_size, config.hidden_size, bias=False)


class BailingMoeGate(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.top_k = config.num_experts_per_tok
        self.num_experts = config.num_experts
        self.norm_topk_prob = config.norm_topk_prob
        self.weight = nn.Parameter(torch.empty((config.num_experts, config.hidden_size)))


class BailingMoeSparseMoeBlock(nn.Module):
    """Parameters of the MoE block; `forward` is the MoE operator boundary of SURVEY.md §8(b):
    forward(hidden_states, image_mask=None, audio_mask=None) -> (y, (router_logits, topk_idx))."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        self.num_experts_per_tok = config.num_experts_per_tok
        self.experts = nn.ModuleList([BailingMoeMLP(config, config.moe_intermediate_size)
                                      for _ in range(config.num_experts)])
        self.multi_gate = config.multi_gate
        if self.multi_gate:
            self.image_gate = BailingMoeGate(config)
            self.audio_gate = BailingMoeGate(config)
        self.gate = BailingMoeGate(config)
        if config.num_shared_experts is not None and config.num_shared_experts > 0:
            self.shared_experts = BailingMoeMLP(config, config.moe_intermediate_size * config.num_shared_experts)
        self._pk = None
        # expert parallelism: this rank keeps experts [ep_rank * E / ep_size, (ep_rank + 1) * E / ep_size)
        self.ep_group, self.ep_rank, self.ep_size = None, 0, 1
        self.ep_mode, self._a2a, self.ep_peer = "allreduce", None, None
        self._slab = None
        _packs.watch(self, self._reset_packs)

    def _reset_packs(self) -> None:
        self._pk, self._a2a = None, None

    # below this many tokens per rank the all-to-all exchange costs more than it saves (decode: B <= 3 rows)
    A2A_MIN_TOKENS_PER_RANK = 4

    def set_expert_parallel(self, group, rank: int, size: int, mode: str = "allreduce", peer=None) -> None:
        """Rank `rank` of `size` keeps experts [rank E / size, (rank + 1) E / size).  Exchange modes:
        "dispatch"  (the one that scales) every rank runs its OWN rows; dispatch and combine both go through NVLink peer
                    memory inside the kernels of csrc/ep.cu (ep.PeerDispatch) — no NCCL, graph-capturable;
        "peer"      tokens replicated on every rank, combine through peer memory (ep.PeerExchange, <= 8 rows);
        "allreduce" tokens replicated, NCCL all-reduce of the fp32 partial sums;
        "alltoall"  prefill-sized inputs also shard the TOKENS of the block with NCCL all-to-alls (ep.py)."""
        if self.config.num_experts % size != 0:
            raise ValueError("num_experts must be divisible by the expert-parallel world size")
        if mode not in ("allreduce", "alltoall", "peer", "dispatch"):
            raise ValueError("mode must be 'allreduce', 'alltoall', 'peer' or 'dispatch'")
        self.ep_group, self.ep_rank, self.ep_size = (group if size > 1 else None), rank, size
        self.ep_mode = mode
        self.ep_peer = peer if (mode in ("peer", "dispatch") and size > 1) else None
        if mode in ("peer", "dispatch") and size > 1 and peer is None:
            raise ValueError(f"mode '{mode}' needs its exchange area (ming_univision_b200.ep)")
        self._a2a = None
        self._pk = None

    def materialize_experts(self, device, ep_rank: int = 0, ep_size: int = 1, dtype=BF16) -> None:
        """Gives the LOCAL routed experts their storage as views into two contiguous slabs — Wgu [E_local, 2I, D]
        (gate rows then up rows) and Wd [E_local, D, I], the layout the expert kernels stream — and leaves the experts of
        other ranks without storage (meta tensors).  `load_state_dict` / in-place initialisation then write straight into
        the slabs: no per-expert tensors next to a stacked copy, and an expert-parallel rank holds only its
        E / G experts.  Call on a module whose expert parameters are still on the meta device (or to be discarded)."""
        cfg = self.config
        E, I, D = cfg.num_experts, cfg.moe_intermediate_size, cfg.hidden_size
        if E % ep_size != 0:
            raise ValueError("num_experts must be divisible by the expert-parallel world size")
        n_local = E // ep_size
        e0 = ep_rank * n_local
        Wgu = torch.empty((n_local, 2 * I, D), dtype=dtype, device=device)
        Wd = torch.empty((n_local, D, I), dtype=dtype, device=device)
        for e, ex in enumerate(self.experts):
            if e0 <= e < e0 + n_local:
                j = e - e0
                ex.gate_proj.weight = nn.Parameter(Wgu[j, :I], requires_grad=False)
                ex.up_proj.weight = nn.Parameter(Wgu[j, I:], requires_grad=False)
                ex.down_proj.weight = nn.Parameter(Wd[j], requires_grad=False)
            else:
                for lin in (ex.gate_proj, ex.up_proj, ex.down_proj):
                    lin.weight = nn.Parameter(torch.empty(lin.weight.shape, dtype=dtype, device="meta"),
                                              requires_grad=False)
        self._slab = (Wgu, Wd, e0, n_local)
        self._pk = None

    def _pack(self):
        dev = self.gate.weight.device
        if self._pk is None or self._pk["dev"] != dev:
            d = lambda t: t.detach().to(device=dev, dtype=BF16).contiguous()  # noqa: E731
            pk = {"dev": dev, "gate": d(self.gate.weight)}
            if self.multi_gate:
                pk["image_gate"] = d(self.image_gate.weight)
            # per-expert slabs: Wgu[e] = [gate_proj; up_proj] ([2I, D]), Wd[e] = down_proj ([D, I])
            n_local = len(self.experts) // self.ep_size
            pk["e_begin"] = self.ep_rank * n_local
            slab = getattr(self, "_slab", None)
            if slab is not None and slab[2] == pk["e_begin"] and slab[3] == n_local and slab[0].device == dev \
                    and slab[0].dtype == BF16 and \
                    self.experts[pk["e_begin"]].gate_proj.weight.data_ptr() == slab[0].data_ptr():
                pk["Wgu"], pk["Wd"] = slab[0], slab[1]  # the parameters ARE the slabs (materialize_experts)
            else:
                local = list(self.experts)[pk["e_begin"]:pk["e_begin"] + n_local]
                pk["Wgu"] = torch.stack([torch.cat([d(e.gate_proj.weight), d(e.up_proj.weight)], dim=0)
                                         for e in local]).contiguous()
                pk["Wd"] = torch.stack([d(e.down_proj.weight) for e in local]).contiguous()
            if hasattr(self, "shared_experts"):
                s = self.shared_experts
                pk["s12"] = torch.cat([d(s.gate_proj.weight), d(s.up_proj.weight)], dim=0).contiguous()
                pk["s12p"], _, hp = ops.pack_swiglu(pk["s12"], None)
                pk["s3"] = d(s.down_proj.weight)
                pk["s3p"] = ops.pad_cols(pk["s3"], hp)
            self._pk = pk
        return self._pk

    def _apply(self, fn, *a, **k):
        self._reset_packs()
        self._slab = None  # .to() / .bfloat16() re-allocate the parameters: they no longer alias the slabs
        _packs.bump()
        return super()._apply(fn, *a, **k)

    @torch.no_grad()
    def _run(self, x2d: torch.Tensor, residual: Optional[torch.Tensor], image_mask: Optional[torch.Tensor]):
        """x2d [T, D] (post-attention-norm) -> (residual + moe(x) [T, D], router logits [T, E], topk idx [T, k])."""
        pk = self._pack()
        cfg = self.config
        if (self.ep_group is not None and self.ep_mode == "alltoall"
                and x2d.shape[0] >= self.A2A_MIN_TOKENS_PER_RANK * self.ep_size):
            return self._run_alltoall(pk, x2d, residual, image_mask), None, None
        logits = _dense(x2d, pk["gate"])
        logits_img, im = None, None
        if self.multi_gate and image_mask is not None:
            logits_img = _dense(x2d, pk["image_gate"])
            im = image_mask.reshape(-1).to(torch.uint8).contiguous()
        idx, w = ops.router_topk(logits, cfg.num_experts_per_tok, cfg.num_experts_per_tok > 1 and cfg.norm_topk_prob,
                                 logits_img, im)
        dispatch = self.ep_mode == "dispatch" and self.ep_size > 1
        if dispatch and x2d.shape[0] <= self.ep_peer.t_max:
            # push this rank's rows to the expert owners FIRST: the shared-expert GEMMs below then run while the rows cross
            # NVLink and while slower ranks catch up
            ops.ep_dispatch(self.ep_peer, x2d, idx, w)
        shared = None
        if "s12" in pk:
            if x2d.shape[0] <= 8:
                shared = ops.gemv(ops.gemv(x2d, pk["s12"], None, epi=ops.EPI_SWIGLU), pk["s3"])
            else:
                shared = ops.linear(ops.linear(x2d, pk["s12p"], None, epi=ops.EPI_SWIGLU), pk["s3p"])
        if dispatch:
            # data parallel x expert parallel: these are THIS rank's rows; dispatch + combine over peer memory (csrc/ep.cu)
            pd, T = self.ep_peer, x2d.shape[0]
            n_local = pk["Wgu"].shape[0]
            ys = []
            for c0 in range(0, T, pd.t_max):  # every rank makes the same calls (same T everywhere)
                c1 = min(T, c0 + pd.t_max)
                if T > pd.t_max:
                    ops.ep_dispatch(pd, x2d[c0:c1], idx[c0:c1], w[c0:c1])
                out_pairs, pair_row = ops.ep_compute(pd, c1 - c0, pk["Wgu"], pk["Wd"], pk["e_begin"], cfg.num_experts)
                ops.ep_combine(pd, c1 - c0, out_pairs, pair_row, pk["e_begin"], n_local)
                ys.append(ops.ep_finalize(pd, c1 - c0, None if shared is None else shared[c0:c1],
                                          None if residual is None else residual[c0:c1]))
            return (ys[0] if len(ys) == 1 else torch.cat(ys, dim=0)), logits, idx
        y = ops.moe_experts(x2d, idx, w, pk["Wgu"], pk["Wd"], shared, residual, pk["e_begin"], self.ep_group,
                            self.ep_peer if self.ep_mode == "peer" else None)
        return y, logits, idx

    @torch.no_grad()
    def _run_alltoall(self, pk, x2d: torch.Tensor, residual: Optional[torch.Tensor], image_mask: Optional[torch.Tensor]):
        """Token- and expert-sharded MoE block: this rank routes tokens [t0, t1), sends every
        (token, slot) row to the rank owning its expert (all-to-all), runs the grouped tcgen05 expert GEMMs on what it
        receives, returns the rows (all-to-all), combines its tokens in fp32 in slot order (:632-638) with the shared
        expert and the residual, and all-gathers the token slices.  Returns residual + moe(x) for ALL T tokens."""
        from .ep import ExpertParallelAllToAll, token_slice

        if self._a2a is None:
            self._a2a = ExpertParallelAllToAll(self.ep_group)
        a2a, cfg = self._a2a, self.config
        G, r = self.ep_size, self.ep_rank
        T, D = x2d.shape
        k, E = cfg.num_experts_per_tok, cfg.num_experts
        E_local = E // G
        t0, t1, Tc = token_slice(T, G, r)
        n_loc = t1 - t0
        dev = x2d.device
        y_loc = torch.zeros((Tc, D), dtype=BF16, device=dev)
        if n_loc > 0:
            x_loc = x2d[t0:t1].contiguous()
            logits = _dense(x_loc, pk["gate"])
            logits_img, im = None, None
            if self.multi_gate and image_mask is not None:
                logits_img = _dense(x_loc, pk["image_gate"])
                im = image_mask.reshape(-1)[t0:t1].to(torch.uint8).contiguous()
            idx, w = ops.router_topk(logits, k, k > 1 and cfg.norm_topk_prob, logits_img, im)
            shared = None
            if "s12" in pk:
                if n_loc <= 8:
                    shared = ops.gemv(ops.gemv(x_loc, pk["s12"], None, epi=ops.EPI_SWIGLU), pk["s3"])
                else:
                    shared = ops.linear(ops.linear(x_loc, pk["s12p"], None, epi=ops.EPI_SWIGLU), pk["s3p"])
            # send buffer ordered by destination rank (bucket = expert // E_local), no padding
            pair_row, row_token, _, _, n_send, counts = ops.moe_plan(idx, G, 0, div=E_local, granule=1,
                                                                      want_counts=True)
            send_rows = ops.gather_rows(x_loc, row_token, n_send)
            send_ids = torch.empty((n_send,), dtype=torch.int32, device=dev)
            send_ids[pair_row.long()] = idx.reshape(-1)  # pure indexing: expert id of every row in send order
        else:
            counts = torch.zeros((G,), dtype=torch.int32, device=dev)
            send_rows = torch.empty((0, D), dtype=BF16, device=dev)
            send_ids = torch.empty((0,), dtype=torch.int32, device=dev)
        send_splits, recv_splits = a2a.exchange_counts(counts)
        recv_rows, recv_ids = a2a.dispatch(send_rows, send_ids, send_splits, recv_splits)
        n_recv = recv_rows.shape[0]
        if n_recv > 0:
            # every received row is one pair for one LOCAL expert: tile plan with k = 1
            pr2, rt2, te2, meta2, mr2 = ops.moe_plan(recv_ids.view(n_recv, 1), E_local, pk["e_begin"])
            xg = ops.gather_rows(recv_rows, rt2, mr2, meta2)
            out_recv = ops.gather_rows(ops.moe_grouped_ffn(xg, pk["Wgu"], pk["Wd"], te2, meta2), pr2, n_recv)
        else:
            out_recv = recv_rows
        back = a2a.combine(out_recv, send_splits, recv_splits)
        if n_loc > 0:
            res_loc = None if residual is None else residual[t0:t1].contiguous()
            ops.moe_combine(back, w, shared, res_loc, y_loc[:n_loc], pair_row)
        return a2a.all_gather_rows(y_loc)[:T]

    @torch.no_grad()
    def forward(self, hidden_states, image_mask=None, audio_mask=None):
        """modeling_bailing_moe.py:556-606 (audio routing is out of scope)."""
        if audio_mask is not None:
            raise NotImplementedError("audio routing is outside the continuous-visual-token path")
        B, S, D = hidden_states.shape
        x2d = hidden_states.reshape(B * S, D).contiguous()
        y, logits, idx = self._run(x2d, None, image_mask)
        if logits is not None and self.multi_gate and image_mask is not None:
            # the reference returns the image gate's logits at image positions (:578-581); row selection only
            logits = torch.where(image_mask.reshape(-1, 1).to(torch.bool), _dense(x2d, self._pack()["image_gate"]), logits)
        if logits is None:  # token-sharded all-to-all mode: every rank routed only its slice; redo the (cheap) router here
            pk, cfg = self._pack(), self.config
            logits = _dense(x2d, pk["gate"])
            li, im = None, None
            if self.multi_gate and image_mask is not None:
                li, im = _dense(x2d, pk["image_gate"]), image_mask.reshape(-1).to(torch.uint8).contiguous()
            idx, _ = ops.router_topk(logits, cfg.num_experts_per_tok,
                                     cfg.num_experts_per_tok > 1 and cfg.norm_topk_prob, li, im)
            if li is not None:
                logits = torch.where(image_mask.reshape(-1, 1).to(torch.bool), li, logits)
        return y.view(B, S, D), (logits.view(B, S, -1), idx.view(B, S, -1).long())


class BailingMoeAttention(nn.Module):
    def __init__(self, config, layer_idx=None):
        super().__init__()
        self.layer_idx = layer_idx
        hd = config.head_dim
        self.query_key_value = nn.Linear(config.hidden_size, (config.num_attention_heads + 2 * config.num_key_value_heads) * hd,
                                         bias=config.use_qkv_bias)
        self.dense = nn.Linear(config.num_attention_heads * hd, config.hidden_size, bias=config.use_bias)


class BailingMoeDecoderLayer(nn.Module):
    def __init__(self, config, layer_idx):
        super().__init__()
        if config.num_experts is None or layer_idx < config.first_k_dense_replace:
            raise NotImplementedError("dense (non-MoE) decoder layers are not on the Ming-UniVision path "
                                      "(first_k_dense_replace = 0, mingunivision/config.json:41)")
        self.attention = BailingMoeAttention(config, layer_idx)
        self.mlp = BailingMoeSparseMoeBlock(config)
        self.input_layernorm = BailingMoeRMSNorm(config.hidden_size, eps=config.rms_norm_eps)
        self.post_attention_layernorm = BailingMoeRMSNorm(config.hidden_size, eps=config.rms_norm_eps)


def _dense(x2d: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, **kw) -> torch.Tensor:
    """nn.Linear on the streaming kernel (<= 8 rows) or the tcgen05 GEMM."""
    if x2d.shape[0] <= 8:
        return ops.gemv(x2d, w, bias, **kw)
    return ops.linear(x2d, w, bias, **kw)


class BailingKVCache:
    """Static KV cache: per layer K, V of shape [Bmax, Hkv, Tmax, hd] (bf16).  Stands in for the HF DynamicCache the
    reference threads through `past_key_values` (cache .repeat / trim of generate_image, :1891-1902 / :1954-1962,
    become a row copy / a batch counter)."""

    def __init__(self, config, max_batch: int, max_len: int, device):
        L, Hkv, hd = config.num_hidden_layers, config.num_key_value_heads, config.head_dim
        self.k = [torch.zeros((max_batch, Hkv, max_len, hd), dtype=BF16, device=device) for _ in range(L)]
        self.v = [torch.zeros((max_batch, Hkv, max_len, hd), dtype=BF16, device=device) for _ in range(L)]
        self.seq_len = 0
        self.batch = 1
        self.max_len = max_len
        self.max_batch = max_batch

    def get_seq_length(self, layer_idx: int = 0) -> int:
        return self.seq_len

    def repeat_rows(self, B: int) -> None:
        if B > self.max_batch:
            raise ValueError("KV cache allocated for fewer rows")
        for k, v in zip(self.k, self.v):
            for b in range(self.batch, B):
                k[b, :, :self.seq_len].copy_(k[0, :, :self.seq_len])
                v[b, :, :self.seq_len].copy_(v[0, :, :self.seq_len])
        self.batch = B

    def trim_rows(self) -> None:
        self.batch = 1

    def expand_groups(self, G: int, B: int) -> None:
        """Rows 0 .. G-1 (the cond prefills of G samples generated together) -> rows g*B + b, b < B: every sample's CFG
        rows start as copies of its cond row (the batched form of repeat_rows)."""
        if G * B > self.max_batch:
            raise ValueError("KV cache allocated for fewer rows")
        if B > 1:
            for k, v in zip(self.k, self.v):
                for g in range(G - 1, -1, -1):  # downwards: a destination row never holds a source that is still needed
                    for b in range(B - 1, -1, -1):
                        if g * B + b != g:
                            k[g * B + b, :, :self.seq_len].copy_(k[g, :, :self.seq_len])
                            v[g * B + b, :, :self.seq_len].copy_(v[g, :, :self.seq_len])
        self.batch = G * B

    def trim_groups(self, G: int, B: int) -> None:
        """Back to one (cond) row per sample: row g*B -> row g (the batched form of trim_rows)."""
        if B > 1:
            for k, v in zip(self.k, self.v):
                for g in range(1, G):
                    k[g, :, :self.seq_len].copy_(k[g * B, :, :self.seq_len])
                    v[g, :, :self.seq_len].copy_(v[g * B, :, :self.seq_len])
        self.batch = G

    def grow(self, min_len: int) -> None:
        """Re-allocates the cache for at least `min_len` tokens (doubling), keeping what it holds — the stand-in for the
        reference's DynamicCache growing with every torch.cat (later editing rounds outgrow any fixed size).  Graph
        workspaces are keyed on (cache, max_len), so the decode graphs are re-captured on the new buffers."""
        new_len = max(min_len, 2 * self.max_len)
        for i in range(len(self.k)):
            for buf in (self.k, self.v):
                old = buf[i]
                new = torch.zeros((old.shape[0], old.shape[1], new_len, old.shape[3]), dtype=old.dtype, device=old.device)
                new[:, :, :self.seq_len].copy_(old[:, :, :self.seq_len])
                buf[i] = new
        self.max_len = new_len


class BailingMoeModel(nn.Module):
    def __init__(self, config: BailingMoeConfig):
        super().__init__()
        self.config = config
        self.vocab_size = config.vocab_size
        self.word_embeddings = nn.Embedding(config.vocab_size, config.hidden_size)
        self.layers = nn.ModuleList([BailingMoeDecoderLayer(config, i) for i in range(config.num_hidden_layers)])
        self.norm = BailingMoeRMSNorm(config.hidden_size, eps=config.rms_norm_eps)
        self._pk = None
        _packs.watch(self, self._reset_packs)

    def _reset_packs(self) -> None:
        self._pk = None

    def _apply(self, fn, *a, **k):
        self._reset_packs()
        _packs.bump()
        return super()._apply(fn, *a, **k)

    def _pack(self):
        dev = self.norm.weight.device
        if dev.type != "cuda":
            raise RuntimeError("BailingMoeModel (B200-native) runs on CUDA only; there is no CPU fallback")
        if self._pk is None or self._pk["dev"] != dev:
            d = lambda t: None if t is None else t.detach().to(device=dev, dtype=BF16).contiguous()  # noqa: E731
            layers = []
            for lyr in self.layers:
                a = lyr.attention
                layers.append(dict(ln1=d(lyr.input_layernorm.weight), ln2=d(lyr.post_attention_layernorm.weight),
                                   qkv_w=d(a.query_key_value.weight), qkv_b=d(a.query_key_value.bias),
                                   dense_w=d(a.dense.weight), dense_b=d(a.dense.bias)))
            self._pk = dict(dev=dev, layers=layers, norm=d(self.norm.weight), emb=d(self.word_embeddings.weight))
        return self._pk

    def set_expert_parallel(self, group=None, mode: str = "allreduce", t_max: int = 256, peer=None) -> None:
        """Shards the routed experts of every layer over the ranks of `group` (default: the world group); attention,
        gates, shared experts and norms stay replicated.  One process per GPU.  Modes: see
        BailingMoeSparseMoeBlock.set_expert_parallel; "dispatch" ("data parallel x expert parallel": every rank runs its
        own rows, `t_max` = most rows per rank and MoE call) and "peer" exchange through NVLink peer memory without any
        NCCL call, so the CUDA-graph fast paths stay on.  `peer`: an existing exchange area to reuse."""
        import torch.distributed as dist

        size = dist.get_world_size(group) if dist.is_initialized() else 1
        rank = dist.get_rank(group) if dist.is_initialized() else 0
        if peer is not None:  # (tests: virtual ranks on one device)
            size, rank = peer.size, peer.rank
        elif mode == "peer" and size > 1:  # one exchange area for all layers (stream-ordered calls, epoch protocol)
            from .ep import PeerExchange

            peer = PeerExchange(group, self.config.hidden_size, self.norm.weight.device)
        elif mode == "dispatch" and size > 1:
            from .ep import PeerDispatch

            peer = PeerDispatch(group, self.config.hidden_size, self.config.num_experts_per_tok, t_max,
                                self.norm.weight.device)
        grp = group if group is not None else (dist.group.WORLD if size > 1 and dist.is_initialized() else None)
        for lyr in self.layers:
            lyr.mlp.set_expert_parallel(grp, rank, size, mode, peer)
        self.ep_size, self.ep_mode, self.ep_group, self.ep_peer = size, mode, grp, peer
        # replicated-token modes need identical inputs (hence identical RF noise) on every rank; "dispatch" does not
        self.ep_tokens_replicated = mode != "dispatch"
        self.ep_graphable = size == 1 or mode in ("peer", "dispatch")

    def check_expert_parallel(self) -> None:
        """Raises if a bounded wait of the peer-memory exchange expired since the areas were created (a peer crashed or
        diverged: csrc/ep.cu records an error code instead of hanging or trapping).  Reads 8 bytes of the local exchange
        area, i.e. synchronises the stream: called once per request by the entry points, not per layer."""
        peer = getattr(self, "ep_peer", None)
        if peer is not None and hasattr(peer, "check"):
            peer.check()

    def embed(self, input_ids: torch.Tensor) -> torch.Tensor:
        """word_embeddings lookup (row gather of the packed bf16 table; pure indexing, no arithmetic)."""
        return self._pack()["emb"][input_ids]

    @torch.no_grad()
    def forward_tokens(self, inputs_embeds: torch.Tensor, position_ids: torch.Tensor, cache: BailingKVCache,
                       key_mask: Optional[torch.Tensor] = None, image_mask: Optional[torch.Tensor] = None,
                       t_dev: Optional[torch.Tensor] = None) -> torch.Tensor:
        """BailingMoeModel.forward (:1391-1540) for this path's two regimes.
        inputs_embeds [B, S, D]; position_ids int [B, S]; the S new tokens are appended at cache slots seq_len..;
        S == 1: cached decode of B rows, `key_mask` int32 [B, >= seq_len+1] marks attendable slots (2-D padding mask of
        the CFG rows); S > 1: causal prefill behind whatever the cache already holds (all-ones mask; the first round
        starts from an empty cache, later rounds append their prompt).  Returns final-norm hidden [B, S, D].
        With `t_dev` (device int32 scalar = current cache length, S must be 1) the call has fixed shapes and reads the
        position from device memory, so it can be captured in a CUDA graph; the caller advances cache.seq_len."""
        pk = self._pack()
        cfg = self.config
        B, S, D = inputs_embeds.shape
        H, hd = cfg.num_attention_heads, cfg.head_dim
        graph_mode = t_dev is not None
        if graph_mode and S != 1:
            raise ValueError("t_dev is for the single-token decode step")
        t0 = 0 if graph_mode else cache.seq_len
        if not graph_mode and t0 + S > cache.max_len:
            cache.grow(t0 + S)
        if S > 1 and key_mask is not None and bool((key_mask[:, :t0 + S] == 0).any()):
            raise NotImplementedError("multi-token forward needs an all-ones key mask (causal prefill; a later round's "
                                      "prompt appended behind the cached context)")
        h = inputs_embeds.to(BF16).reshape(B * S, D).contiguous().clone()
        # 3-D multimodal RoPE (rope_scaling.type == "3D", :413-425 / :463-469) is config-gated: position_ids [3, B, S]
        mrope = position_ids.dim() == 3
        if mrope:
            rs = cfg.rope_scaling
            # (transformers 5 normalises rope_scaling into a dict with a `rope_type` key; the reference reads `type`)
            is3d = isinstance(rs, dict) and "3D" in (rs.get("type"), rs.get("rope_type"))
            if not is3d or tuple(position_ids.shape) != (3, B, S):
                raise ValueError("3-D position_ids [3, B, S] need config.rope_scaling = {'type': '3D', ...}")
            pos = position_ids.reshape(3, B * S).to(torch.int32).contiguous()
        else:
            pos = position_ids.reshape(-1).to(torch.int32).contiguous()
        eps = cfg.rms_norm_eps
        im = None if image_mask is None else image_mask.reshape(-1)
        for li, (lp, lyr) in enumerate(zip(pk["layers"], self.layers)):
            # decode: RMSNorm fused into the staging of the qkv streaming GEMM — measured on the full-size edit round:
            # 771.6 -> 737.9 ms with 3 rows, no change with 6 (the rows take three register rounds there).  Every
            # decode-sized input takes the SAME form, so that requests generated together (6 rows) get bit for bit the
            # result they get alone (3 rows): the fused statistics are reduced in another order than rmsnorm_kernel's
            if B * S <= 8 and D <= 4096 and (FUSED_NORM or _FUSED_ENV is None):
                qkv = ops.gemv_norm(h, lp["qkv_w"], lp["qkv_b"], norm="rms", gamma=lp["ln1"], eps=eps)
            else:
                qkv = _dense(ops.rmsnorm(h, lp["ln1"], eps), lp["qkv_w"], lp["qkv_b"])
            # rows 0..B-1 of the [Bmax, Hkv, Tmax, hd] cache are a contiguous prefix the kernels index directly
            if mrope:
                q = ops.rope3d_kv_append(qkv, pos, cache.k[li], cache.v[li], B, S, H, t0, cfg.rope_theta,
                                         (16, 24, 24), t_dev)  # the reference's hard-wired sections (:463)
            else:
                q = ops.rope_kv_append(qkv, pos, cache.k[li], cache.v[li], B, S, H, t0, cfg.rope_theta, t_dev)
            if S == 1:
                a = ops.attn_decode_gqa(q, cache.k[li], cache.v[li], key_mask, H, t0 + 1, t_dev)
            else:
                a = ops.attn_prefill_gqa(q, cache.k[li], cache.v[li], B, S, H, t0)
            _dense(a, lp["dense_w"], lp["dense_b"], epi=ops.EPI_RESIDUAL, residual=h, out=h)
            x = ops.rmsnorm(h, lp["ln2"], eps)
            h, _, _ = lyr.mlp._run(x, h, im)
        if not graph_mode:
            cache.seq_len = t0 + S
        return ops.rmsnorm(h, pk["norm"], eps).view(B, S, D)


class BailingMoeForCausalLM(nn.Module):
    """Parameters under the reference's keys (model.*, lm_head.*, vis_head.*, diffloss.*) and the image-generation
    entry points of :1543-1965."""

    def __init__(self, config: BailingMoeConfig):
        super().__init__()
        self.config = config
        self.model = BailingMoeModel(config)
        self.vocab_size = config.vocab_size
        self.norm_head = config.norm_head
        self.lm_head = nn.Linear(config.hidden_size, config.vocab_size, bias=False)
        self.vis_head = None
        self.diffloss = None
        self.num_generated_images = 0
        self._pk = None
        self.use_cuda_graph = True
        # The CFG rows of a generation carry IDENTICAL latents (shared noise, one combined velocity: diff_loss_rf_swiglu.py
        # :118-122, :149-171), so the reference's semantic-decoder step, linear_proj and pixel decoder compute the same
        # row B times (:1939-1964).  With this package's own callbacks those three stages run on ONE row and the result is
        # broadcast — bit-identical (every kernel on that path is row-independent), (B-1)/B of their traffic saved.
        self.dedupe_cfg_rows = True
        self._gen_ws = {}
        self._txt_ws = {}
        _packs.watch(self, self._reset_packs)
        for p in self.parameters():
            p.requires_grad_(False)

    def _reset_packs(self) -> None:
        self._pk = None
        self._gen_ws = {}
        self._txt_ws = {}

    def _apply(self, fn, *a, **k):
        self._reset_packs()
        _packs.bump()
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        """Accepts the reference's checkpoints unchanged: its per-layer `rotary_emb.inv_freq` buffers
        (BailingMoeRotaryEmbeddingLegacy, modeling_bailing_moe.py:213-237) are derived from `rope_theta` here (the RoPE
        kernel recomputes them), so those keys are dropped instead of being reported as unexpected."""
        sd = {k: v for k, v in state_dict.items() if not k.endswith("rotary_emb.inv_freq")}
        return super().load_state_dict(sd, strict=strict, **kw)

    def setup_vishead_diffloss(self, diffloss_w=3072, diffloss_d=12, num_sampling_steps="16",
                               gen_method="flow_matching_swiglu-4", hidden_size=2048, vis_head_arch="linear2-norm",
                               image_emb_dim_for_gen=32):
        """modeling_bailing_moe.py:1559-1584."""
        assert vis_head_arch == "linear2-norm"
        assert gen_method.startswith("flow_matching_swiglu-")
        self.vis_head = nn.Sequential(nn.Linear(hidden_size, diffloss_w), nn.LayerNorm(diffloss_w, eps=1e-6))
        self.diffloss = RectifiedFlowLoss(target_channels=image_emb_dim_for_gen, z_channels=diffloss_w,
                                          width=diffloss_w, depth=diffloss_d, num_sampling_steps=num_sampling_steps,
                                          mlp_mult=int(gen_method.split("-")[1]), grad_checkpointing=False)
        for p in self.parameters():
            p.requires_grad_(False)
        self._pk = None

    def reset_image_gen_status(self):
        self.num_generated_images = 0

    def get_input_embeddings(self):
        return self.model.word_embeddings

    def _pack(self):
        dev = self.lm_head.weight.device
        if self._pk is None or self._pk["dev"] != dev:
            d = lambda t: t.detach().to(device=dev, dtype=BF16).contiguous()  # noqa: E731
            pk = dict(dev=dev, lm_head=d(self.lm_head.weight))
            if self.vis_head is not None:
                pk.update(vh_w=d(self.vis_head[0].weight), vh_b=d(self.vis_head[0].bias),
                          vh_g=d(self.vis_head[1].weight), vh_beta=d(self.vis_head[1].bias))
            self._pk = pk
        return self._pk

    def new_cache(self, max_len: int, max_batch: int = 3) -> BailingKVCache:
        return BailingKVCache(self.config, max_batch, max_len, self.lm_head.weight.device)

    @torch.no_grad()
    def compute_logit(self, hidden_states: torch.Tensor) -> torch.Tensor:
        """:1604-1620 + the `.float()` of :1785/:1817: bf16 lm_head GEMM, logits returned in fp32."""
        if self.norm_head:
            raise NotImplementedError("norm_head is False on the Ming-UniVision path (config.json)")
        pk = self._pack()
        x = hidden_states.reshape(-1, hidden_states.shape[-1]).to(BF16).contiguous()
        if x.shape[0] <= 8:
            f32 = torch.empty((x.shape[0], self.vocab_size), dtype=torch.float32, device=x.device)
            ops.gemv(x, pk["lm_head"], None, out_f32=f32)
            return f32.view(*hidden_states.shape[:-1], -1)
        return ops.linear(x, pk["lm_head"], None).float().view(*hidden_states.shape[:-1], -1)

    @torch.no_grad()
    def compute_vis_z(self, hidden_last: torch.Tensor) -> torch.Tensor:
        """z = vis_head(h) = LayerNorm(Linear(h)) (:1571-1574, :1657-1662); h [B, D] -> z [B, Z] bf16."""
        pk = self._pack()
        y = _dense(hidden_last.contiguous(), pk["vh_w"], pk["vh_b"])
        return ops.layernorm(y, pk["vh_g"], pk["vh_beta"], 1e-6)

    @torch.no_grad()
    def forward_for_image_generation_inner(self, inputs_embeds, attention_mask, position_ids, past_key_values,
                                           image_gen_temperature=1.0, image_gen_text_cfg=3.0, image_gen_image_cfg=1.1,
                                           noise=None, groups: int = 1, **kwargs):
        """:1622-1673: one LLM step on the given embeddings, z = vis_head(last hidden), latent = diffloss.sample(z).
        attention_mask: int32 [B, >= cache_len + S] key mask (see BailingMoeModel.forward_tokens)."""
        hidden = self.model.forward_tokens(inputs_embeds, position_ids, past_key_values, key_mask=attention_mask)
        z = self.compute_vis_z(hidden[:, -1])
        x = self.diffloss.sample(z, temperature=image_gen_temperature, text_cfg=image_gen_text_cfg,
                                 image_cfg=image_gen_image_cfg, noise=noise, groups=groups)
        return x.unsqueeze(1), hidden

    @torch.no_grad()
    def generate_image(self, input_embeds, past_key_values: BailingKVCache, attention_mask, uncond_attention_mask,
                       text_uncond_attention_mask, latent_to_sem_func, linear_proj, sem_to_pix_func,
                       image_gen_text_cfg=3.0, image_gen_image_cfg=1.1, image_gen_temperature=1.0, noises=None):
        """:1844-1965.  CFG rows are batch rows (B = 2: cond + uncond; B = 3: + text-uncond); the KV cache of the cond
        prefill is replicated to the rows and trimmed back to row 0 afterwards.  As in the reference, the CFG scales
        handed to the sampler are the hard-wired defaults 3.0 / 1.1: `generate_image` passes them under the wrong
        keyword names so they never reach `diffloss.sample` — only the temperature propagates.
        `noises` (test hook): sequence of [G, C] tensors replacing the per-token torch.randn draw.

        Extension (the reference asserts one sequence, :1865): G > 1 independent requests generated
        TOGETHER — input_embeds [G, 1, D], masks [G, n], the cache holding the G cond prefills in rows 0 .. G-1.  Sample g
        owns rows g*B .. g*B + B-1; every weight-streaming kernel of the step (LLM, RF head, semantic decoder) then serves
        G*B <= 8 rows per pass over the weights.  Returns images [G, 3, H, W] (G = 1: [B, ...] as the reference)."""
        cfg = self.config
        dev = input_embeds.device
        G = attention_mask.shape[0]
        attention_mask = attention_mask.to(torch.int32)
        rows = [attention_mask]
        if uncond_attention_mask is not None:
            uncond_attention_mask = uncond_attention_mask.to(torch.int32)
            n_c, n_u = attention_mask.shape[1], uncond_attention_mask.shape[1]
            if n_u < n_c:
                uncond_attention_mask = torch.cat((uncond_attention_mask, attention_mask[:, n_u:]), dim=1)
            rows.append(uncond_attention_mask)
        if text_uncond_attention_mask is not None and int(text_uncond_attention_mask.sum()) > 0:
            text_uncond_attention_mask = text_uncond_attention_mask.to(torch.int32)
            n_c, n_u = attention_mask.shape[1], text_uncond_attention_mask.shape[1]
            if n_u < n_c:
                text_uncond_attention_mask = torch.cat((text_uncond_attention_mask, attention_mask[:, n_u:]), dim=1)
            if int((text_uncond_attention_mask == uncond_attention_mask).sum()) != uncond_attention_mask.numel():
                rows.append(text_uncond_attention_mask)
        B = len(rows)
        attention_mask = torch.stack(rows, dim=1).reshape(G * B, -1)  # row g*B + b = mask b of sample g
        R = G * B
        if R > 8:
            raise ValueError(f"{G} requests x {B} CFG rows: at most 8 rows per generation step")
        n_tok = cfg.num_image_tokens_for_gen
        cache = past_key_values
        if cache.batch != G:
            raise ValueError(f"the KV cache holds {cache.batch} prefilled rows, the masks describe {G} requests")
        if B > 1:
            input_embeds = input_embeds.repeat_interleave(B, dim=0)
            cache.expand_groups(G, B)
        if self.use_cuda_graph and getattr(self.model, "ep_graphable", True) and \
                self._graphable(latent_to_sem_func, linear_proj):
            return self._generate_image_graphed(input_embeds, cache, attention_mask, B, G, n_tok, latent_to_sem_func,
                                                linear_proj, sem_to_pix_func, image_gen_temperature, noises)
        # key mask buffer for the whole generation: prompt part now, one more "1" column per generated token
        t_now = attention_mask.shape[1]
        mask = torch.ones((R, t_now + n_tok + 1), dtype=torch.int32, device=dev)
        mask[:, :t_now] = attention_mask.to(dev)
        pos0 = (attention_mask.long().cumsum(-1) - 1)[:, -1:].to(dev)  # position of the current token per row
        output_tokens, sem_cache, hidden = [], None, None
        one_row = self.dedupe_cfg_rows and B > 1 and self._graphable(latent_to_sem_func, linear_proj)
        for token_idx in range(n_tok + 1):
            position_ids = (pos0 + token_idx).to(torch.int32)
            latent, hidden = self.forward_for_image_generation_inner(
                inputs_embeds=input_embeds, attention_mask=mask, position_ids=position_ids, past_key_values=cache,
                image_gen_temperature=image_gen_temperature, groups=G,
                noise=self._draw_noise(dev, G) if noises is None else noises[token_idx].to(dev))
            if token_idx < n_tok:
                feat = latent_to_sem_func(latent[0::B].contiguous() if one_row else latent, past_key_values=sem_cache)
                sem_cache = feat["past_key_values"]
                output_token = feat["x_norm_patchtokens"]
                output_tokens.append(output_token)
                input_embeds = linear_proj(output_token)
                if one_row:
                    input_embeds = input_embeds.repeat_interleave(B, dim=0)
        cache.trim_groups(G, B)
        final_mask = mask[:, :t_now + n_tok]
        image_tensor = sem_to_pix_func(torch.cat(output_tokens, dim=1))
        if G == 1 and one_row:
            image_tensor = image_tensor.expand(B, *image_tensor.shape[1:])
        elif G > 1 and not one_row:
            image_tensor = image_tensor[0::B]
        return image_tensor, hidden, final_mask

    def _draw_noise(self, dev, groups: int = 1) -> torch.Tensor:
        """The per-token torch.randn(1, C) of RectifiedFlowLoss.sample (diff_loss_rf_swiglu.py:118).  Expert-parallel
        modes that REPLICATE the tokens on every rank need the same draw everywhere (the ranks' partial expert sums are
        added): rank 0 draws, the others receive it, so differently seeded processes cannot diverge silently."""
        noise = torch.randn(groups, self.diffloss.in_channels, device=dev)
        if getattr(self.model, "ep_size", 1) > 1 and getattr(self.model, "ep_tokens_replicated", True):
            import torch.distributed as dist

            dist.broadcast(noise, src=dist.get_global_rank(self.model.ep_group, 0) if self.model.ep_group is not None
                           else 0, group=self.model.ep_group)
        return noise

    # ---------------------------------------------------------------------------------------------------------------
    # Greedy text decoding (HF GenerationMixin.generate with do_sample = false, mingunivision/config.json:30, as driven by
    # MingUniVisionForConditionalGeneration.generate, modeling_bailingmm.py:249-270): one CUDA-graph replay per token
    # ---------------------------------------------------------------------------------------------------------------
    def _text_step(self, ws) -> None:
        """hidden(last token) -> lm_head logits (fp32) -> argmax -> embedding of the chosen token -> one cached LLM step,
        on static buffers with the cache position read from device memory."""
        logits = self.compute_logit(ws["last"])
        ws["tok"].copy_(ops.argmax_rows(logits.reshape(1, -1)))
        ws["embeds"].copy_(self.model.embed(ws["tok"].view(1, 1).long()))
        hidden = self.model.forward_tokens(ws["embeds"], ws["pos"], ws["cache"], key_mask=None, t_dev=ws["t_llm"])
        ws["last"].copy_(hidden[:, -1])
        ws["t_llm"].add_(1)
        ws["pos"].add_(1)

    @torch.no_grad()
    def greedy_decode(self, last_hidden: torch.Tensor, cache: "BailingKVCache", max_new_tokens: int,
                      stop_ids=()) -> list:
        """Greedy continuation after a prefill: `last_hidden` [1, D] is the final-norm hidden state of the last prompt
        token, `cache` holds the prompt.  Returns the new token ids (stops after emitting any id in `stop_ids`; the
        per-token stop check reads the token on the host, as HF generate does).  The cache ends up holding the prompt
        plus every emitted token but the last (HF generate never feeds the final token back either)."""
        dev = last_hidden.device
        if cache.batch != 1:
            raise ValueError("text decoding runs on one row")
        if cache.seq_len + max_new_tokens > cache.max_len:
            cache.grow(cache.seq_len + max_new_tokens)
        use_graph = self.use_cuda_graph and getattr(self.model, "ep_graphable", True)
        key = (id(cache), cache.max_len, _packs.epoch())
        ws = self._txt_ws.get(key) if use_graph else None
        if ws is None:
            D = self.config.hidden_size
            ws = dict(cache=cache, last=torch.zeros((1, D), dtype=BF16, device=dev),
                      embeds=torch.zeros((1, 1, D), dtype=BF16, device=dev),
                      tok=torch.zeros((1,), dtype=torch.int32, device=dev),
                      pos=torch.zeros((1, 1), dtype=torch.int32, device=dev),
                      t_llm=torch.zeros((1,), dtype=torch.int32, device=dev), graph=None)
            if use_graph:
                self._txt_ws = {key: ws}

        def reset_state():
            ws["last"].copy_(last_hidden.reshape(1, -1).to(BF16))
            ws["pos"].fill_(cache.seq_len)
            ws["t_llm"].fill_(cache.seq_len)

        reset_state()
        if use_graph and ws["graph"] is None:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):  # warm-up outside capture; its cache write is overwritten by the real run
                self._text_step(ws)
            torch.cuda.current_stream().wait_stream(side)
            reset_state()
            g = torch.cuda.CUDAGraph()
            l0 = _lib.launch_count()
            with torch.cuda.graph(g):
                self._text_step(ws)
            ws["graph"], ws["kernels"] = g, _lib.launch_count() - l0
            reset_state()
        out = []
        stop = set(int(t) for t in stop_ids)
        for _ in range(max_new_tokens):
            if use_graph:
                ws["graph"].replay()
                _lib.count_replay(ws["kernels"])
            else:
                self._text_step(ws)
            t = int(ws["tok"].item())
            out.append(t)
            cache.seq_len += 1  # the step that chose `t` also fed it through the model
            if t in stop:
                break
        if out:
            cache.seq_len -= 1  # as with HF generate, the last emitted token is not part of the cached context
        return out

    # ---------------------------------------------------------------------------------------------------------------
    # CUDA-graph fast path of the AR visual-token loop: one replay per generated token
    # ---------------------------------------------------------------------------------------------------------------
    @staticmethod
    def _graphable(latent_to_sem_func, linear_proj) -> bool:
        """The whole token step can be captured when the callbacks are this package's own modules (fixed shapes, no
        host sync): MingTok.forward_feature_decoder and LinearProj."""
        from .mingtok.modeling_mingtok import MingTok
        from .modeling_bailingmm import LinearProj

        vision = getattr(latent_to_sem_func, "__self__", None)
        return isinstance(vision, MingTok) and getattr(latent_to_sem_func, "__name__", "") == "forward_feature_decoder" \
            and isinstance(linear_proj, LinearProj)

    def _token_step(self, ws) -> None:
        """LLM step -> vis_head -> RF sampler (hard-wired CFG 3.0 / 1.1) -> semantic-decoder step -> linear_proj, all on
        static buffers and device-side positions (forward_for_image_generation_inner + the loop body of generate_image).
        R = G * B rows: G requests generated together, B CFG rows each (row g*B + b)."""
        B, G = ws["B"], ws["G"]
        hidden = self.model.forward_tokens(ws["embeds"], ws["pos"], ws["cache"], key_mask=ws["mask"], t_dev=ws["t_llm"])
        ws["hidden"].copy_(hidden)
        z = self.compute_vis_z(hidden[:, -1])
        C = ws["noise"].shape[1]
        ws["x"].view(G, B, C).copy_((ws["noise"] * ws["temperature"]).unsqueeze(1).expand(G, B, C))
        self.diffloss._sample_body(self.diffloss._pack(), z, ws["x"], 3.0, 1.1, B if B in (2, 3) else 1)
        if ws["sem_rows"] == G * B:   # one semantic-decoder row per LLM row (CFG rows not de-duplicated)
            lat = ws["x"]
        elif G == 1:
            lat = ws["x"][0:1]
        else:                          # row 0 of every sample (the CFG rows of a sample carry identical latents)
            ws["lat"].copy_(ws["x"].view(G, B, C)[:, 0])
            lat = ws["lat"]
        feat = ws["vision"]._decode_step(lat, ws["sem_cache"], 0, ws["sem_cache"].t_dev)
        ws["feats"].index_copy_(1, ws["t_idx"], feat.unsqueeze(1))
        e = ws["linear_proj"](feat.unsqueeze(1))
        D = e.shape[-1]
        if ws["sem_rows"] == G * B:
            ws["embeds"].copy_(e)
        else:
            ws["embeds"].view(G, B, 1, D).copy_(e.view(G, 1, 1, D).expand(G, B, 1, D))
        ws["t_llm"].add_(1)
        ws["sem_cache"].t_dev.add_(1)
        ws["t_idx"].add_(1)
        ws["pos"].add_(1)

    def _generate_image_graphed(self, input_embeds, cache, attention_mask, B, G, n_tok, latent_to_sem_func, linear_proj,
                                sem_to_pix_func, temperature, noises):
        dev = input_embeds.device
        vision = latent_to_sem_func.__self__
        t_now = attention_mask.shape[1]
        R = G * B
        if cache.seq_len != t_now - 1:
            raise ValueError("KV cache / attention mask length mismatch")
        if cache.seq_len + n_tok + 1 > cache.max_len:
            cache.grow(cache.seq_len + n_tok + 1)
        S = G if self.dedupe_cfg_rows else R   # semantic-decoder rows
        key = (B, G, S, id(cache), id(vision), id(linear_proj), n_tok, float(temperature), cache.max_len, _packs.epoch())
        ws = self._gen_ws.get(key)
        if ws is None:
            C, D, F = self.diffloss.in_channels, self.config.hidden_size, vision.feature_dim
            ws = dict(B=B, G=G, sem_rows=S, temperature=float(temperature), cache=cache, vision=vision,
                      linear_proj=linear_proj,
                      embeds=torch.zeros((R, 1, D), dtype=BF16, device=dev),
                      hidden=torch.zeros((R, 1, D), dtype=BF16, device=dev),
                      noise=torch.zeros((G, C), dtype=torch.float32, device=dev),
                      x=torch.zeros((R, C), dtype=torch.float32, device=dev),
                      lat=torch.zeros((G, C), dtype=torch.float32, device=dev),
                      feats=torch.zeros((S, n_tok + 1, F), dtype=BF16, device=dev),
                      pos=torch.zeros((R, 1), dtype=torch.int32, device=dev),
                      mask=torch.ones((R, cache.max_len), dtype=torch.int32, device=dev),
                      t_llm=torch.zeros((1,), dtype=torch.int32, device=dev),
                      t_idx=torch.zeros((1,), dtype=torch.int64, device=dev),
                      sem_cache=vision.new_decode_cache(S, n_tok + 8), graph=None)
            self._gen_ws = {key: ws}  # one workspace at a time (a new cache / batch size re-captures)

        def reset_state():
            ws["embeds"].copy_(input_embeds.to(BF16))
            ws["mask"].fill_(1)
            ws["mask"][:, :t_now] = attention_mask.to(dev)
            ws["pos"].copy_((attention_mask.long().cumsum(-1) - 1)[:, -1:].to(torch.int32))
            ws["t_llm"].fill_(cache.seq_len)
            ws["t_idx"].zero_()
            ws["sem_cache"].reset()

        reset_state()
        if ws["graph"] is None:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):  # warm-up outside capture; its cache writes are overwritten by the real run
                self._token_step(ws)
            torch.cuda.current_stream().wait_stream(side)
            reset_state()
            g = torch.cuda.CUDAGraph()
            l0 = _lib.launch_count()
            with torch.cuda.graph(g):
                self._token_step(ws)
            ws["graph"], ws["kernels"] = g, _lib.launch_count() - l0
            reset_state()
        for token_idx in range(n_tok + 1):
            # RNG stays on the host side as in the reference (torch.randn(1, C) per token, diff_loss_rf_swiglu.py:118)
            ws["noise"].copy_(self._draw_noise(dev, G) if noises is None else noises[token_idx].to(dev))
            ws["graph"].replay()
            _lib.count_replay(ws["kernels"])
        cache.seq_len += n_tok + 1
        cache.trim_groups(G, B)
        final_mask = ws["mask"][:, :t_now + n_tok].clone()
        image_tensor = sem_to_pix_func(ws["feats"][:, :n_tok])
        if G == 1 and S != R:
            image_tensor = image_tensor.expand(B, *image_tensor.shape[1:])
        elif G > 1 and S == R:
            image_tensor = image_tensor[0::B]
        self._last_gen_latent_hidden = ws["hidden"]
        return image_tensor, ws["hidden"].clone(), final_mask
