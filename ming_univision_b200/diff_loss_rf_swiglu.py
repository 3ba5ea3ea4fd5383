"""B200-native rectified-flow SwiGLU head: same class surface and state_dict keys as the reference's
``mingunivision/diff_loss_rf_swiglu.py`` (RectifiedFlowLoss :76-181, SimpleMLPAdaLN :295-385, ResBlock :242-272,
FinalLayer :275-292, TimestepEmbedder :188-239), executed by the weight-streaming kernels of libmingb200.so.

The head is HBM-bound (M = CFG rows <= 3): 1.285 B parameters = 2.57 GB of bf16 weights per network evaluation,
16 evaluations per visual token.  Two restructurings keep the reference's arithmetic but cut the traffic:
  * time_embed(t) takes only `num_sampling_steps` values -> a [steps, W] table computed once at pack time;
    cond_embed(z) is step-invariant -> once per token (SURVEY.md §7 "hard parts" (c));
  * every adaLN_modulation Linear depends on the step only through SiLU(t_emb[s] + c), so all 13 of them are applied
    to the 16*B conditioning rows in ONE GEMM per token: their 358 M parameters (28 % of the head) are streamed once per
    token instead of 16 times (41.1 GB -> 29.7 GB of weight traffic per token).
The Euler loop (16 x [input_proj, 12 x (adaLN-modulate, w12+SwiGLU, w3+gated residual), final layer, CFG+Euler]) is
captured into a CUDA graph per batch size, so the ~650 launches per token cost no host time.
"""
from __future__ import annotations

import math
import os

import torch
import torch.nn as nn

from . import _lib, _packs, ops

BF16 = torch.bfloat16


def _swiglu_hidden(hidden_features: int) -> int:
    return (int(hidden_features * 2 / 3) + 7) // 8 * 8  # SwiGLUFFNFused, diff_loss_rf_swiglu.py:66


# ---------------------------------------------------------------------------------------------------------------
# parameter containers (names == reference attribute names; no torch forward is ever called)
# ---------------------------------------------------------------------------------------------------------------
class _SwiGLU(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.w12 = nn.Linear(dim, 2 * hidden, bias=True)
        self.w3 = nn.Linear(hidden, dim, bias=True)


class TimestepEmbedder(nn.Module):
    def __init__(self, hidden_size, frequency_embedding_size=256):
        super().__init__()
        self.mlp = nn.Sequential(nn.Linear(frequency_embedding_size, hidden_size, bias=True), nn.SiLU(),
                                 nn.Linear(hidden_size, hidden_size, bias=True))
        self.frequency_embedding_size = frequency_embedding_size


class ResBlock(nn.Module):
    def __init__(self, channels, mlp_mult=1.0):
        super().__init__()
        self.channels = channels
        self.in_ln = nn.LayerNorm(channels, eps=1e-6)
        self.mlp = _SwiGLU(channels, _swiglu_hidden(int(channels * mlp_mult)))
        self.adaLN_modulation = nn.Sequential(nn.SiLU(), nn.Linear(channels, 3 * channels, bias=True))


class FinalLayer(nn.Module):
    def __init__(self, model_channels, out_channels):
        super().__init__()
        self.norm_final = nn.LayerNorm(model_channels, elementwise_affine=False, eps=1e-6)
        self.linear = nn.Linear(model_channels, out_channels, bias=True)
        self.adaLN_modulation = nn.Sequential(nn.SiLU(), nn.Linear(model_channels, 2 * model_channels, bias=True))


class SimpleMLPAdaLN(nn.Module):
    def __init__(self, in_channels, model_channels, out_channels, z_channels, num_res_blocks, mlp_mult=1.0,
                 grad_checkpointing=False):
        super().__init__()
        self.in_channels, self.model_channels, self.out_channels = in_channels, model_channels, out_channels
        self.num_res_blocks = num_res_blocks
        self.time_embed = TimestepEmbedder(model_channels)
        self.cond_embed = nn.Linear(z_channels, model_channels)
        self.input_proj = nn.Linear(in_channels, model_channels)
        self.res_blocks = nn.ModuleList([ResBlock(model_channels, mlp_mult) for _ in range(num_res_blocks)])
        self.final_layer = FinalLayer(model_channels, out_channels)


# The streaming kernel can normalise its input rows itself (ops.gemv_norm: LN + adaLN modulation computed by all 256
# threads of every CTA while its first weight batch is in flight).  Measured on B200 (tools/bench_rf.py): with B = 2 rows
# the fused form wins (8.23 -> 7.89 ms per RF sample: 208 row-kernel launches and their dependency gaps disappear); with
# B = 3 the rows no longer fit one register round and the separate row kernel, which overlaps the previous GEMM's tail
# under PDL, is as fast (8.88 vs 9.01 ms).  MB_FUSED_NORM = 1 / 0 forces either form; default: fused for B <= 2.
_FUSED_ENV = os.environ.get("MB_FUSED_NORM")
FUSED_NORM = _FUSED_ENV == "1"  # (LLM decode step: fused RMSNorm only when forced)


# MB_RF_FUSED = 0 keeps the launch-per-layer path (A/B timing, tests); default: the persistent sampler kernel whenever
# the shape fits it (B <= 3 rows, widths multiples of 1024 — the default head; the tiny test heads take the layer path).
def _use_fused(pk, n_rows: int) -> bool:
    if os.environ.get("MB_RF_FUSED", "1") == "0":
        return False
    H = pk.blocks[0][2].shape[0] // 2
    return bool(_lib.load().mb_rf_fused_supported(n_rows, pk.W, H, pk.C))


# bench.py: when set to a list, every launch of the persistent sampler kernel is bracketed by CUDA events on the launching
# stream and (algorithmic weight bytes, start, end) is appended (eager path only; events cannot be captured into a graph)
FUSED_PROFILE: list | None = None


def _fuse_adaln(n_rows: int) -> bool:
    return _FUSED_ENV == "1" if _FUSED_ENV in ("0", "1") else n_rows <= 2


class _PackedRF:
    def __init__(self, m: "RectifiedFlowLoss", device):
        d = lambda t: t.detach().to(device=device, dtype=BF16).contiguous()  # noqa: E731
        net = m.net
        self.device = device
        self.W = net.model_channels
        self.C = net.in_channels
        self.steps = m.num_sampling_steps
        self.cond_w, self.cond_b = d(net.cond_embed.weight), d(net.cond_embed.bias)
        self.in_w, self.in_b = d(net.input_proj.weight), d(net.input_proj.bias)
        self.blocks = []
        ada_w, ada_b = [], []
        for rb in net.res_blocks:
            self.blocks.append((d(rb.in_ln.weight), d(rb.in_ln.bias), d(rb.mlp.w12.weight), d(rb.mlp.w12.bias),
                                d(rb.mlp.w3.weight), d(rb.mlp.w3.bias)))
            ada_w.append(rb.adaLN_modulation[1].weight)
            ada_b.append(rb.adaLN_modulation[1].bias)
        ada_w.append(net.final_layer.adaLN_modulation[1].weight)
        ada_b.append(net.final_layer.adaLN_modulation[1].bias)
        # all adaLN modulation Linears stacked: [depth*3W + 2W, W]
        self.ada_w = d(torch.cat([w.detach() for w in ada_w], dim=0))
        self.ada_b = d(torch.cat([b.detach() for b in ada_b], dim=0))
        self.fin_w, self.fin_b = d(net.final_layer.linear.weight), d(net.final_layer.linear.bias)
        # time-embedding table for the fixed schedule t = 1, 1 - 1/steps, ... (sample(), :134-136; t*1000 at :372)
        t = torch.linspace(1.0, 0.0, self.steps + 1, device=device)[:-1] * 1000
        half = net.time_embed.frequency_embedding_size // 2
        freqs = torch.exp(-math.log(10000) * torch.arange(0, half, dtype=torch.float32, device=device) / half)
        args = t[:, None].float() * freqs[None]
        tfreq = torch.cat([torch.cos(args), torch.sin(args)], dim=-1).to(BF16).contiguous()  # cos first (:231)
        w0, b0 = d(net.time_embed.mlp[0].weight), d(net.time_embed.mlp[0].bias)
        w2, b2 = d(net.time_embed.mlp[2].weight), d(net.time_embed.mlp[2].bias)
        rows = []
        for s0 in range(0, self.steps, 8):  # the streaming kernel takes <= 8 rows per call
            h = ops.gemv(tfreq[s0:s0 + 8], w0, b0, epi=ops.EPI_SILU)
            rows.append(ops.gemv(h, w2, b2))
        self.temb = torch.cat(rows, dim=0).contiguous()  # [steps, W] bf16
        self._fused = None

    def fused(self):
        """Per-CTA stage-ordered copies of the w12 / w3 matrices + the device pointer table and scratch of the persistent
        sampler kernel (mb_rf_sample_fused); built on first use."""
        if self._fused is None:
            lib = _lib.load()
            n_cta = lib.mb_num_sms()
            dev, W = self.device, self.W
            H = self.blocks[0][2].shape[0] // 2
            keep, table = [], []
            s = torch.cuda.current_stream().cuda_stream
            for (lnw, lnb, w12, b12, w3, b3) in self.blocks:
                w12p, w3p = torch.empty_like(w12), torch.empty_like(w3)
                _lib.check(lib.mb_rf_pack_weights(w12.data_ptr(), w12.shape[0], w12.shape[1], 1, n_cta, w12p.data_ptr(), s),
                           "mb_rf_pack_weights")
                _lib.check(lib.mb_rf_pack_weights(w3.data_ptr(), w3.shape[0], w3.shape[1], 0, n_cta, w3p.data_ptr(), s),
                           "mb_rf_pack_weights")
                keep += [w12p, w3p]
                table.append([w12p.data_ptr(), b12.data_ptr(), w3p.data_ptr(), b3.data_ptr(), lnw.data_ptr(),
                              lnb.data_ptr()])
            self._fused = dict(n_cta=n_cta, H=H, keep=keep, table=torch.tensor(table, dtype=torch.int64, device=dev),
                               h=torch.zeros((8, W), dtype=BF16, device=dev), hid=torch.zeros((8, H), dtype=BF16, device=dev),
                               v=torch.zeros((8, self.C), dtype=BF16, device=dev),
                               bar=torch.zeros((16,), dtype=torch.int32, device=dev))
        return self._fused


class RectifiedFlowLoss(nn.Module):
    """Drop-in for diff_loss_rf_swiglu.RectifiedFlowLoss (inference: `sample`)."""

    def __init__(self, target_channels, z_channels, depth, width, num_sampling_steps, mlp_mult=1.0,
                 grad_checkpointing=False):
        super().__init__()
        self.in_channels = target_channels
        self.num_sampling_steps = int(num_sampling_steps) if isinstance(num_sampling_steps, str) else num_sampling_steps
        self.net = SimpleMLPAdaLN(in_channels=target_channels, model_channels=width, out_channels=target_channels,
                                  z_channels=z_channels, num_res_blocks=depth, mlp_mult=mlp_mult,
                                  grad_checkpointing=grad_checkpointing)
        self.t_sample_strategy = "uniform"
        self.use_cuda_graph = True
        self._packed: _PackedRF | None = None
        self._graphs: dict = {}
        _packs.watch(self, self._reset_packs)
        for p in self.parameters():
            p.requires_grad_(False)

    def _reset_packs(self) -> None:
        self._packed, self._graphs = None, {}

    def set_t_sample_strategy(self, strategy="uniform"):
        self.t_sample_strategy = strategy

    def _apply(self, fn, *args, **kwargs):
        self._reset_packs()
        _packs.bump()
        return super()._apply(fn, *args, **kwargs)

    def _pack(self) -> _PackedRF:
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("RectifiedFlowLoss (B200-native) runs on CUDA only; there is no CPU fallback")
        if self._packed is None or self._packed.device != dev:
            self._packed = _PackedRF(self, dev)
            self._graphs = {}
        return self._packed

    # -- the sampler body (graph-capturable: fixed shapes, no host sync) ---------------------------------------
    def _sample_body(self, pk: _PackedRF, z_bf16: torch.Tensor, x_f32: torch.Tensor, text_cfg: float,
                     image_cfg: float, cfg_rows: int | None = None) -> None:
        """z_bf16 [B, Z], x_f32 [B, C] (noise in, sample out).  The B rows are B / cfg_rows independent samples (images
        generated together) of `cfg_rows` adjacent CFG rows each — the weights stream once for all of them."""
        B, W, depth = z_bf16.shape[0], pk.W, len(pk.blocks)
        cfg_rows = B if cfg_rows is None else cfg_rows
        fuse = _fuse_adaln(B)
        c = ops.gemv(z_bf16, pk.cond_w, pk.cond_b)                       # cond_embed(z), once per token (:374)
        sy = ops.silu_add_rows(pk.temb, c)                               # SiLU(t_emb[s] + c) for every step
        mod = ops.linear(sy, pk.ada_w, pk.ada_b)                         # all adaLN modulations, [steps*B, depth*3W+2W]
        if _use_fused(pk, B):
            # the 16-step Euler loop as ONE persistent weight-streaming kernel (csrc/rf_fused.cu)
            f = pk.fused()
            lib = _lib.load()
            prof = FUSED_PROFILE
            if prof is not None:
                ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ev0.record()
            _lib.check(lib.mb_rf_sample_fused(f["table"].data_ptr(), pk.in_w.data_ptr(), pk.in_b.data_ptr(),
                                              pk.fin_w.data_ptr(), pk.fin_b.data_ptr(), mod.data_ptr(), mod.stride(0),
                                              x_f32.data_ptr(), f["h"].data_ptr(), f["hid"].data_ptr(), f["v"].data_ptr(),
                                              f["bar"].data_ptr(), B, cfg_rows, W, f["H"], pk.C, depth, pk.steps,
                                              float(text_cfg), float(image_cfg), f["n_cta"],
                                              torch.cuda.current_stream().cuda_stream), "mb_rf_sample_fused")
            if prof is not None:
                ev1.record()
                prof.append((pk.steps * depth * 3 * f["H"] * W * 2, ev0, ev1))
            return
        x_bf16 = ops.affine(x_f32, 1.0, 0.0)
        dt = 1.0 / pk.steps
        for s in range(pk.steps):
            ms = mod[s * B:(s + 1) * B]
            h = ops.gemv(x_bf16, pk.in_w, pk.in_b)                       # input_proj (:371)
            for i, (lnw, lnb, w12, b12, w3, b3) in enumerate(pk.blocks):  # ResBlock.forward (:268-272)
                o = i * 3 * W
                if fuse:  # LN + adaLN modulation fused into the staging of the w12 streaming GEMM
                    hid = ops.gemv_norm(h, w12, b12, norm="adaln", gamma=lnw, beta=lnb, shift=ms[:, o:o + W],
                                        scale=ms[:, o + W:o + 2 * W], epi=ops.EPI_SWIGLU)
                else:
                    a = ops.adaln_modulate(h, lnw, lnb, ms[:, o:o + W], ms[:, o + W:o + 2 * W])
                    hid = ops.gemv(a, w12, b12, epi=ops.EPI_SWIGLU)
                ops.gemv(hid, w3, b3, epi=ops.EPI_GATED, residual=h, gate=ms[:, o + 2 * W:o + 3 * W], out=h)
            o = depth * 3 * W                                            # FinalLayer.forward (:288-292)
            if fuse:
                v = ops.gemv_norm(h, pk.fin_w, pk.fin_b, norm="adaln", shift=ms[:, o:o + W], scale=ms[:, o + W:o + 2 * W])
            else:
                a = ops.adaln_modulate(h, None, None, ms[:, o:o + W], ms[:, o + W:o + 2 * W])
                v = ops.gemv(a, pk.fin_w, pk.fin_b)
            ops.rf_euler_step(x_f32, x_bf16, v, dt, text_cfg, image_cfg, cfg_rows)  # CFG combine + Euler (:145-179)

    @torch.no_grad()
    def sample(self, z, temperature=1.0, text_cfg=1.0, image_cfg=1.0, cfg_renorm_type=None,
               time_shifting_factor=None, noise=None, groups: int = 1):
        """diff_loss_rf_swiglu.py:103-181.  z: [B, z_channels] (fp32 or bf16, CUDA).  Returns x: [B, C] fp32.
        `noise` (optional, test hook) replaces the torch.randn draw: [groups, C] if text_cfg != 1 else [B, C].
        `groups` > 1 (extension; the reference samples one image at a time): the B rows are `groups` independent samples of
        B / groups adjacent CFG rows — several images share every pass over the weights."""
        if cfg_renorm_type is not None or time_shifting_factor:
            raise NotImplementedError("cfg_renorm_type / time_shifting_factor are always None on the reference path "
                                      "(modeling_bailing_moe.py:1859-1860)")
        pk = self._pack()
        B = z.shape[0]
        if groups < 1 or B % groups != 0 or B > 8:
            raise ValueError(f"{B} rows are not {groups} whole samples (<= 8 rows per call)")
        rows = B // groups
        device = z.device
        # RNG stays on the host side exactly as in the reference (:117-122) so the Philox stream matches
        if noise is None:
            noise = torch.randn(groups if text_cfg != 1.0 else B, self.in_channels, device=device)
        x0 = (noise.repeat_interleave(rows, dim=0) if text_cfg != 1.0 else noise) * temperature
        # the reference selects the CFG combine by the ROW COUNT of the sample (b_num == 3 / == 2, :146-171), any other
        # count integrates every row on its own (:172-173)
        cfg_rows = rows if rows in (2, 3) else 1
        key = (B, cfg_rows, float(text_cfg), float(image_cfg), _packs.epoch())
        if not self.use_cuda_graph:
            x = x0.float().contiguous().clone()
            self._sample_body(pk, ops.affine(z.reshape(B, -1), 1.0, 0.0), x, text_cfg, image_cfg, cfg_rows)
            return x
        if key not in self._graphs:
            z_buf = torch.zeros((B, z.shape[-1]), dtype=BF16, device=device)
            x_buf = torch.zeros((B, self.in_channels), dtype=torch.float32, device=device)
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):  # warm-up outside capture (sets kernel attributes, fills the allocator)
                self._sample_body(pk, z_buf, x_buf, text_cfg, image_cfg, cfg_rows)
            torch.cuda.current_stream().wait_stream(side)
            g = torch.cuda.CUDAGraph()
            l0 = _lib.launch_count()
            with torch.cuda.graph(g):
                self._sample_body(pk, z_buf, x_buf, text_cfg, image_cfg, cfg_rows)
            self._graphs[key] = (g, z_buf, x_buf, _lib.launch_count() - l0)
        g, z_buf, x_buf, n_kernels = self._graphs[key]
        z_buf.copy_(z.reshape(B, -1))
        x_buf.copy_(x0)
        g.replay()
        _lib.count_replay(n_kernels)
        return x_buf.clone()
