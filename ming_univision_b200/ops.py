"""torch.Tensor -> raw pointer adapters over the C ABI (include/mingb200.h).

PyTorch is used for device memory and the current CUDA stream only; every function here launches hand-written
sm_100a kernels from libmingb200.so.  No function has an eager-PyTorch fallback.
"""
from __future__ import annotations

import os

import torch

from . import _lib

EPI_BIAS, EPI_GELU, EPI_SWIGLU, EPI_RESIDUAL = 0, 1, 2, 3
BF16 = torch.bfloat16

# When set to a list, linear() brackets every GEMM launch with CUDA events on the launching stream and appends
# ((M, N, K, epi), flops, start_event, end_event) — used by bench.py for the roofline of the dominant kernel.
PROFILE: list | None = None
# When set to a list, linear() also appends a zero-argument closure that re-issues exactly the same launch (same
# buffers, shapes, epilogue): bench.py replays the step's GEMM launches back to back between two events, which times
# the dominant kernel without the per-launch event pairs (those serialise the PDL chain and add ~2 us per launch).
REPLAY: list | None = None
# The same for the weight-streaming kernel (gemv / gemv_norm): (closure, algorithmic bytes = N*K*2, (M, N, K, epi)).
GEMV_REPLAY: list | None = None


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: torch.Tensor | None) -> int | None:
    return None if t is None else t.data_ptr()


def _check_bf16(*ts: torch.Tensor | None) -> None:
    for t in ts:
        if t is not None and (t.dtype != BF16 or not t.is_cuda):
            raise TypeError(f"expected a CUDA bfloat16 tensor, got {t.dtype} on {t.device}")


def _rows2d(t: torch.Tensor) -> torch.Tensor:
    """View [..., D] as [rows, D] with a unit inner stride (no copy unless the layout forces one)."""
    t2 = t.reshape(-1, t.shape[-1])
    if t2.stride(-1) != 1:
        t2 = t2.contiguous()
    return t2


def round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


# ---------------------------------------------------------------------------------------------------------------
# GEMM
# ---------------------------------------------------------------------------------------------------------------
def linear(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor | None = None, *, epi: int = EPI_BIAS,
           residual: torch.Tensor | None = None, res_row_mod: int = 0, out: torch.Tensor | None = None,
           out_row_group: int = 0, out_row_pad: int = 0, ln_fold: tuple | None = None,
           stats_out: torch.Tensor | None = None) -> torch.Tensor:
    """out = epilogue(x @ weight.T + bias) on the tcgen05 GEMM.  x: [..., K], weight: [N, K] (nn.Linear layout).
    ln_fold = (stats [M, S, 2] fp32, csum [N] fp32, bias_f32 [N] fp32, eps): LayerNorm of x folded into the epilogue
    (weight must be the gamma-scaled pack, see fold_layernorm); stats_out [M, ceil(N/64), 2] fp32: the RESIDUAL
    epilogue leaves per-row, per-64-column-box (sum, sum of squares) of what it wrote for the next folded GEMM."""
    _check_bf16(x, weight, bias, residual, out)
    lib = _lib.load()
    a = _rows2d(x)
    M, K = a.shape
    N = weight.shape[0]
    if weight.shape[1] != K or weight.stride(1) != 1:
        raise ValueError(f"weight shape {tuple(weight.shape)} incompatible with input K={K}")
    n_out = N // 2 if epi == EPI_SWIGLU else N
    if out is None:
        if out_row_group:
            raise ValueError("out_row_group needs a caller-provided output")
        out2 = torch.empty((M, n_out), dtype=BF16, device=x.device)
        ret = out2.view(*x.shape[:-1], n_out)
    else:
        out2 = _rows2d(out)
        if out2.data_ptr() != out.data_ptr():
            raise ValueError("out must be row-contiguous")
        ret = out
    r2, ldr = None, 0
    if epi == EPI_RESIDUAL:
        if residual is None:
            raise ValueError("EPI_RESIDUAL needs a residual tensor")
        r2 = _rows2d(residual)
        ldr = r2.stride(0)
    prof = PROFILE
    if prof is not None:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
    if ln_fold is None and stats_out is None:
        rc = lib.mb_gemm_bf16(a.data_ptr(), a.stride(0), weight.data_ptr(), weight.stride(0), _ptr(bias),
                              out2.data_ptr(), out2.stride(0), M, N, K, epi, _ptr(r2), ldr, res_row_mod,
                              out_row_group, out_row_pad, _stream())
    else:
        st_in = cs = bf = None
        eps, slots = 0.0, 0
        if ln_fold is not None:
            st_in, cs, bf, eps = ln_fold
            if st_in.dtype != torch.float32 or cs.dtype != torch.float32 or bf.dtype != torch.float32:
                raise TypeError("ln_fold tensors must be fp32")
            if st_in.dim() != 3 or st_in.shape[0] != M or st_in.shape[2] != 2 or not st_in.is_contiguous():
                raise ValueError("ln_fold statistics must be a contiguous [M, S, 2] tensor")
            slots = st_in.shape[1]
        if stats_out is not None and (tuple(stats_out.shape) != (M, (n_out + 63) // 64, 2) or
                                      stats_out.dtype != torch.float32 or not stats_out.is_contiguous()):
            raise ValueError(f"stats_out must be a contiguous fp32 [{M}, {(n_out + 63) // 64}, 2] tensor")
        rc = lib.mb_gemm_bf16_ex(a.data_ptr(), a.stride(0), weight.data_ptr(), weight.stride(0), _ptr(bias),
                                 out2.data_ptr(), out2.stride(0), M, N, K, epi, _ptr(r2), ldr, res_row_mod,
                                 out_row_group, out_row_pad, _ptr(st_in), slots, _ptr(cs), _ptr(bf), float(eps),
                                 _ptr(stats_out), _stream())
    if prof is not None:
        ev1.record()
        prof.append(((M, N, K, epi), 2.0 * M * N * K, ev0, ev1))
    if REPLAY is not None:
        keep = (a, weight, bias, out2, r2, ln_fold, stats_out)  # keeps the buffers alive for the replay
        REPLAY.append((lambda: linear(x, weight, bias, epi=epi, residual=residual, res_row_mod=res_row_mod, out=ret,
                                      out_row_group=out_row_group, out_row_pad=out_row_pad, ln_fold=ln_fold,
                                      stats_out=stats_out), 2.0 * M * N * K, keep))
    _lib.check(rc, "mb_gemm_bf16")
    return ret


def row_stats(x: torch.Tensor) -> torch.Tensor:
    """[rows, D] bf16 -> [rows, 1, 2] fp32 (sum, sum of squares): seeds a chain of LayerNorm-folded GEMMs."""
    _check_bf16(x)
    lib = _lib.load()
    x2 = _rows2d(x)
    st = torch.empty((x2.shape[0], 1, 2), dtype=torch.float32, device=x.device)
    _lib.check(lib.mb_row_stats(x2.data_ptr(), x2.stride(0), st.data_ptr(), x2.shape[0], x2.shape[1], _stream()),
               "mb_row_stats")
    return st


def fold_layernorm(w: torch.Tensor, b: torch.Tensor | None, gamma: torch.Tensor, beta: torch.Tensor):
    """Load-time pack for a LayerNorm folded into the Linear that follows it:
    LN(x) W^T + b = rstd (x W'^T - mean csum) + b'   with W' = bf16(W * gamma), csum = rowsum(W') and
    b' = b + W beta (both fp32).  Returns (W' bf16 [N, K], csum fp32 [N], b' fp32 [N])."""
    wf = w.float()
    wp = (wf * gamma.float()[None, :]).to(BF16).contiguous()
    csum = wp.float().sum(dim=1).contiguous()
    bp = (wf @ beta.float()) + (b.float() if b is not None else 0.0)
    return wp, csum, bp.contiguous()


def pack_swiglu_f32(v12: torch.Tensor, H: int, Hp: int) -> torch.Tensor:
    """The row interleave of pack_swiglu for an fp32 per-row vector [2H] -> [2Hp] (folded csum / bias)."""
    out = torch.zeros((2 * Hp,), dtype=torch.float32, device=v12.device)
    o = out.view(Hp // 128, 2, 128)
    g = torch.zeros((Hp,), dtype=torch.float32, device=v12.device)
    u = torch.zeros((Hp,), dtype=torch.float32, device=v12.device)
    g[:H], u[:H] = v12[:H], v12[H:]
    o[:, 0, :] = g.view(-1, 128)
    o[:, 1, :] = u.view(-1, 128)
    return out


def pack_swiglu(w12: torch.Tensor, b12: torch.Tensor | None) -> tuple[torch.Tensor, torch.Tensor | None, int]:
    """Packs the reference's w12 ([2H, K], x1 rows then x2 rows) for the fused SwiGLU epilogue.
    Returns (packed weight [2*Hp, K], packed bias [2*Hp] or None, Hp)."""
    _check_bf16(w12, b12)
    lib = _lib.load()
    H = w12.shape[0] // 2
    K = w12.shape[1]
    Hp = round_up(H, 128)
    wp = torch.empty((2 * Hp, K), dtype=BF16, device=w12.device)
    _lib.check(lib.mb_pack_swiglu_rows(w12.contiguous().data_ptr(), wp.data_ptr(), H, Hp, K, _stream()),
               "mb_pack_swiglu_rows")
    bp = None
    if b12 is not None:
        bp = torch.empty((2 * Hp,), dtype=BF16, device=w12.device)
        _lib.check(lib.mb_pack_swiglu_rows(b12.contiguous().data_ptr(), bp.data_ptr(), H, Hp, 1, _stream()),
                   "mb_pack_swiglu_rows")
    return wp, bp, Hp


def pad_cols(w: torch.Tensor, Kp: int) -> torch.Tensor:
    """Zero-pads the input dimension of a weight [N, K] to Kp (w3 after the SwiGLU hidden padding)."""
    N, K = w.shape
    if K == Kp:
        return w.contiguous()
    wp = torch.zeros((N, Kp), dtype=w.dtype, device=w.device)
    wp[:, :K] = w
    return wp


# ---------------------------------------------------------------------------------------------------------------
# normalisation / attention
# ---------------------------------------------------------------------------------------------------------------
def layernorm(x: torch.Tensor, gamma: torch.Tensor | None, beta: torch.Tensor | None, eps: float = 1e-6,
              act: int = 0, drop_last_token: bool = False) -> torch.Tensor:
    """LayerNorm over the last dim.  drop_last_token: x is [B, n+1, D] and only the first n tokens of every image are
    normalised and returned densely as [B, n, D] (the decoder's `x_norm[:, :-1]`)."""
    _check_bf16(x, gamma, beta)
    lib = _lib.load()
    if drop_last_token:
        if x.dim() != 3 or not x.is_contiguous():
            raise ValueError("drop_last_token needs a contiguous [B, n+1, D] input")
        B, n1, D = x.shape
        y = torch.empty((B, n1 - 1, D), dtype=BF16, device=x.device)
        rc = lib.mb_layernorm(x.data_ptr(), D, _ptr(gamma), _ptr(beta), y.data_ptr(), D, B * (n1 - 1), D, float(eps),
                              act, n1 - 1, n1 * D, _stream())
        _lib.check(rc, "mb_layernorm")
        return y
    x2 = _rows2d(x)
    y = torch.empty_like(x2, memory_format=torch.contiguous_format)
    rc = lib.mb_layernorm(x2.data_ptr(), x2.stride(0), _ptr(gamma), _ptr(beta), y.data_ptr(), y.stride(0),
                          x2.shape[0], x2.shape[1], float(eps), act, 0, 0, _stream())
    _lib.check(rc, "mb_layernorm")
    return y.view(x.shape)


def set_attn_backend(backend: int) -> None:
    """0 auto / 1 tcgen05 kernel whenever eligible, 2 force the mma.sync kernel (tests, A/B measurements)."""
    rc = _lib.load().mb_attn_set_backend(int(backend))
    if rc != 0:
        raise RuntimeError("mb_attn_set_backend failed")


def attention_hd64(qkv: torch.Tensor, B: int, S: int, H: int, causal: bool) -> torch.Tensor:
    """qkv: [B, S, 3*H*64] packed as (3, H, 64) -> [B, S, H*64]."""
    _check_bf16(qkv)
    lib = _lib.load()
    if not qkv.is_contiguous() or qkv.numel() != B * S * 3 * H * 64:
        raise ValueError("qkv must be contiguous [B, S, 3*H*64]")
    out = torch.empty((B, S, H * 64), dtype=BF16, device=qkv.device)
    rc = lib.mb_attn_hd64(qkv.data_ptr(), out.data_ptr(), B, S, H, 64 ** -0.5, int(causal), _stream())
    _lib.check(rc, "mb_attn_hd64")
    return out


def attention_hd64_decode(qkv: torch.Tensor, kcache: torch.Tensor, vcache: torch.Tensor, t: int,
                          t_dev: torch.Tensor | None = None) -> torch.Tensor:
    """qkv: [B, 3*H*64] for the new token; caches [B, H, Tmax, 64]; appends at t (+ *t_dev) and attends to 0..t."""
    _check_bf16(qkv, kcache, vcache)
    lib = _lib.load()
    B, H, Tmax, hd = kcache.shape
    if hd != 64 or not kcache.is_contiguous() or not vcache.is_contiguous() or not qkv.is_contiguous():
        raise ValueError("caches must be contiguous [B, H, Tmax, 64]")
    out = torch.empty((B, H * 64), dtype=BF16, device=qkv.device)
    rc = lib.mb_attn_hd64_decode(qkv.data_ptr(), kcache.data_ptr(), vcache.data_ptr(), out.data_ptr(), B, H, t, Tmax,
                                 64 ** -0.5, None if t_dev is None else t_dev.data_ptr(), _stream())
    _lib.check(rc, "mb_attn_hd64_decode")
    return out


# ---------------------------------------------------------------------------------------------------------------
# MingTok data movement
# ---------------------------------------------------------------------------------------------------------------
def patchify(img: torch.Tensor, P: int) -> torch.Tensor:
    lib = _lib.load()
    if not img.is_cuda or img.dtype not in (torch.float32, BF16):
        raise TypeError("img must be a CUDA fp32 or bf16 tensor")
    img = img.contiguous()
    B, Cc, Hh, Ww = img.shape
    rows = torch.empty((B * (Hh // P) * (Ww // P), Cc * P * P), dtype=BF16, device=img.device)
    rc = lib.mb_patchify(img.data_ptr(), int(img.dtype == torch.float32), rows.data_ptr(), B, Cc, Hh, Ww, P,
                         _stream())
    _lib.check(rc, "mb_patchify")
    return rows


def fill_cls_row(x: torch.Tensor, cls: torch.Tensor, pos_cls: torch.Tensor) -> None:
    _check_bf16(x, cls, pos_cls)
    lib = _lib.load()
    B, n1, dim = x.shape
    _lib.check(lib.mb_fill_cls_row(x.data_ptr(), cls.data_ptr(), pos_cls.data_ptr(), B, n1, dim, _stream()),
               "mb_fill_cls_row")


def group_mean(x: torch.Tensor, groups: int) -> torch.Tensor:
    _check_bf16(x)
    lib = _lib.load()
    x2 = _rows2d(x)
    out = torch.empty((x2.shape[0], groups), dtype=BF16, device=x.device)
    _lib.check(lib.mb_group_mean(x2.data_ptr(), x2.stride(0), out.data_ptr(), x2.shape[0], x2.shape[1], groups,
                                 _stream()), "mb_group_mean")
    return out.view(*x.shape[:-1], groups)


def affine(x: torch.Tensor, scale: float, shift: float) -> torch.Tensor:
    """bf16(x * scale + shift); x may be fp32 (RF-sampler latents) or bf16."""
    if not x.is_cuda or x.dtype not in (torch.float32, BF16):
        raise TypeError("affine expects a CUDA fp32 or bf16 tensor")
    lib = _lib.load()
    x = x.contiguous()
    y = torch.empty(x.shape, dtype=BF16, device=x.device)
    _lib.check(lib.mb_affine(x.data_ptr(), int(x.dtype == torch.float32), y.data_ptr(), x.numel(), float(scale),
                             float(shift), _stream()), "mb_affine")
    return y


def inproj_repeat(x: torch.Tensor, W: torch.Tensor, b: torch.Tensor | None) -> torch.Tensor:
    _check_bf16(x, W, b)
    lib = _lib.load()
    x2 = _rows2d(x).contiguous()
    dim, in_dim = W.shape
    out = torch.empty((x2.shape[0], dim), dtype=BF16, device=x.device)
    _lib.check(lib.mb_inproj_repeat(x2.data_ptr(), W.contiguous().data_ptr(), _ptr(b), out.data_ptr(), x2.shape[0],
                                    in_dim, dim, _stream()), "mb_inproj_repeat")
    return out.view(*x.shape[:-1], dim)


def pixel_shuffle(x: torch.Tensor, g: int, f: int, Cc: int) -> torch.Tensor:
    _check_bf16(x)
    lib = _lib.load()
    x = x.contiguous()
    B = x.shape[0]
    out = torch.empty((B, g * f * g * f, Cc), dtype=BF16, device=x.device)
    _lib.check(lib.mb_pixel_shuffle(x.data_ptr(), out.data_ptr(), B, g, f, Cc, _stream()), "mb_pixel_shuffle")
    return out


def unpatchify_clamp(x: torch.Tensor, g: int, p: int, out_dtype: torch.dtype = BF16) -> torch.Tensor:
    _check_bf16(x)
    lib = _lib.load()
    x = x.contiguous()
    B = x.shape[0]
    img = torch.empty((B, 3, g * p, g * p), dtype=out_dtype, device=x.device)
    _lib.check(lib.mb_unpatchify_clamp(x.data_ptr(), img.data_ptr(), int(out_dtype == torch.float32), B, g, p,
                                       _stream()), "mb_unpatchify_clamp")
    return img


# ---------------------------------------------------------------------------------------------------------------
# decode regime: weight-streaming skinny GEMM + rectified-flow row helpers
# ---------------------------------------------------------------------------------------------------------------
EPI_SILU, EPI_GATED = 4, 5


def gemv(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor | None = None, *, epi: int = EPI_BIAS,
         residual: torch.Tensor | None = None, gate: torch.Tensor | None = None, out: torch.Tensor | None = None,
         out_f32: torch.Tensor | None = None) -> torch.Tensor:
    """out = epilogue(x @ weight.T + bias) for M <= 8 rows on the HBM-streaming kernel (mb_gemv_bf16).
    EPI_SWIGLU takes the reference w12 layout [2H, K]; EPI_GATED computes residual + gate * (...)."""
    _check_bf16(x, weight, bias, residual, gate, out)
    lib = _lib.load()
    a = _rows2d(x)
    M, K = a.shape
    N = weight.shape[0]
    if weight.shape[1] != K or weight.stride(1) != 1:
        raise ValueError(f"weight shape {tuple(weight.shape)} incompatible with input K={K}")
    n_out = N // 2 if epi == EPI_SWIGLU else N
    if out is None:
        out = torch.empty((M, n_out), dtype=BF16, device=x.device)
    for t in (residual, gate):
        if t is not None and t.stride(-1) != 1:
            raise ValueError("residual / gate must have a unit inner stride")
    rc = lib.mb_gemv_bf16(a.data_ptr(), a.stride(0), weight.data_ptr(), weight.stride(0), _ptr(bias), out.data_ptr(),
                          out.stride(0), M, N, K, epi, _ptr(residual), residual.stride(0) if residual is not None else 0,
                          _ptr(gate), gate.stride(0) if gate is not None else 0, _ptr(out_f32), _stream())
    _lib.check(rc, "mb_gemv_bf16")
    if GEMV_REPLAY is not None:  # bench.py: re-issue the step's streaming launches back to back (algorithmic bytes = N K 2)
        GEMV_REPLAY.append((lambda: gemv(x, weight, bias, epi=epi, residual=residual, gate=gate, out=out,
                                         out_f32=out_f32), 2.0 * N * K, (M, N, K, epi)))
    return out


def gemv_norm(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor | None = None, *, norm: str,
              gamma: torch.Tensor | None = None, beta: torch.Tensor | None = None, shift: torch.Tensor | None = None,
              scale: torch.Tensor | None = None, eps: float = 1e-6, epi: int = EPI_BIAS,
              residual: torch.Tensor | None = None, gate: torch.Tensor | None = None, out: torch.Tensor | None = None,
              out_f32: torch.Tensor | None = None) -> torch.Tensor:
    """gemv() with the input normalisation fused into the kernel (mb_gemv_bf16_norm): norm = "adaln" computes
    (LN(x) * gamma + beta) * (1 + scale) + shift, norm = "rms" computes gamma * (x * rsqrt(mean(x^2) + eps))."""
    _check_bf16(x, weight, bias, residual, gate, out, gamma, beta, shift, scale)
    lib = _lib.load()
    a = _rows2d(x)
    M, K = a.shape
    N = weight.shape[0]
    if weight.shape[1] != K or weight.stride(1) != 1:
        raise ValueError(f"weight shape {tuple(weight.shape)} incompatible with input K={K}")
    n_out = N // 2 if epi == EPI_SWIGLU else N
    if out is None:
        out = torch.empty((M, n_out), dtype=BF16, device=x.device)
    for t in (residual, gate, shift, scale):
        if t is not None and t.stride(-1) != 1:
            raise ValueError("residual / gate / shift / scale must have a unit inner stride")
    code = {"adaln": 1, "rms": 2}[norm]
    rc = lib.mb_gemv_bf16_norm(a.data_ptr(), a.stride(0), weight.data_ptr(), weight.stride(0), _ptr(bias),
                               out.data_ptr(), out.stride(0), M, N, K, epi, _ptr(residual),
                               residual.stride(0) if residual is not None else 0, _ptr(gate),
                               gate.stride(0) if gate is not None else 0, _ptr(out_f32), code, _ptr(gamma), _ptr(beta),
                               _ptr(shift), shift.stride(0) if shift is not None else 0, _ptr(scale),
                               scale.stride(0) if scale is not None else 0, float(eps), _stream())
    _lib.check(rc, "mb_gemv_bf16_norm")
    if GEMV_REPLAY is not None:
        GEMV_REPLAY.append((lambda: gemv_norm(x, weight, bias, norm=norm, gamma=gamma, beta=beta, shift=shift,
                                              scale=scale, eps=eps, epi=epi, residual=residual, gate=gate, out=out,
                                              out_f32=out_f32), 2.0 * N * K, (M, N, K, epi)))
    return out


def adaln_modulate(x: torch.Tensor, gamma: torch.Tensor | None, beta: torch.Tensor | None, shift: torch.Tensor,
                   scale: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    """bf16((LN(x) * gamma + beta) * bf16(1 + scale) + shift) row-wise; shift / scale may be strided row views."""
    _check_bf16(x, gamma, beta, shift, scale)
    lib = _lib.load()
    rows, dim = x.shape
    if shift.stride(-1) != 1 or scale.stride(-1) != 1 or x.stride(-1) != 1:
        raise ValueError("unit inner strides required")
    y = torch.empty((rows, dim), dtype=BF16, device=x.device)
    rc = lib.mb_adaln_modulate(x.data_ptr(), x.stride(0), _ptr(gamma), _ptr(beta), shift.data_ptr(), shift.stride(0),
                               scale.data_ptr(), scale.stride(0), y.data_ptr(), dim, rows, dim, float(eps), _stream())
    _lib.check(rc, "mb_adaln_modulate")
    return y


def silu_add_rows(temb: torch.Tensor, c: torch.Tensor) -> torch.Tensor:
    """[steps, D], [B, D] -> [steps*B, D]: bf16(silu(bf16(temb[s] + c[b]))), row index s*B + b."""
    _check_bf16(temb, c)
    lib = _lib.load()
    steps, dim = temb.shape
    B = c.shape[0]
    out = torch.empty((steps * B, dim), dtype=BF16, device=c.device)
    _lib.check(lib.mb_silu_add_rows(temb.contiguous().data_ptr(), c.contiguous().data_ptr(), out.data_ptr(), steps, B,
                                    dim, _stream()), "mb_silu_add_rows")
    return out


def rf_euler_step(x_f32: torch.Tensor, x_bf16: torch.Tensor, v: torch.Tensor, dt: float, text_cfg: float,
                  image_cfg: float, cfg_rows: int | None = None) -> None:
    """In-place CFG combine + Euler update of the fp32 state x (and its bf16 copy); the B rows are B / cfg_rows
    independent samples of `cfg_rows` adjacent CFG rows (default: one sample)."""
    _check_bf16(x_bf16, v)
    if x_f32.dtype != torch.float32 or not x_f32.is_contiguous() or not v.is_contiguous():
        raise TypeError("x_f32 must be a contiguous fp32 tensor and v contiguous bf16")
    lib = _lib.load()
    B, C = x_f32.shape
    _lib.check(lib.mb_rf_euler_step(x_f32.data_ptr(), x_bf16.data_ptr(), v.data_ptr(), B, int(cfg_rows or B), C,
                                    float(dt), float(text_cfg), float(image_cfg), _stream()), "mb_rf_euler_step")


# ---------------------------------------------------------------------------------------------------------------
# Bailing-MoE AR step
# ---------------------------------------------------------------------------------------------------------------
def _iptr(t: torch.Tensor | None) -> int | None:
    if t is None:
        return None
    if t.dtype != torch.int32 or not t.is_cuda:
        raise TypeError("expected a CUDA int32 tensor")
    return t.data_ptr()


def rmsnorm(x: torch.Tensor, w: torch.Tensor, eps: float) -> torch.Tensor:
    _check_bf16(x, w)
    lib = _lib.load()
    x2 = _rows2d(x)
    y = torch.empty_like(x2, memory_format=torch.contiguous_format)
    _lib.check(lib.mb_rmsnorm(x2.data_ptr(), x2.stride(0), w.data_ptr(), y.data_ptr(), y.stride(0), x2.shape[0],
                              x2.shape[1], float(eps), _stream()), "mb_rmsnorm")
    return y.view(x.shape)


def rope_kv_append(qkv: torch.Tensor, position_ids: torch.Tensor, kcache: torch.Tensor, vcache: torch.Tensor,
                   B: int, S: int, H: int, t: int, rope_theta: float, t_dev: torch.Tensor | None = None) -> torch.Tensor:
    """qkv [B*S, (H+2Hkv)*hd] -> rotated q [B*S, H*hd]; rotated K and V written into the caches at slots t..t+S-1."""
    _check_bf16(qkv, kcache, vcache)
    lib = _lib.load()
    _, Hkv, Tmax, hd = kcache.shape
    q = torch.empty((B * S, H * hd), dtype=BF16, device=qkv.device)
    rc = lib.mb_rope_kv_append(qkv.data_ptr(), _iptr(position_ids), q.data_ptr(), kcache.data_ptr(), vcache.data_ptr(),
                               B, S, H, Hkv, hd, Tmax, _iptr(t_dev), int(t), float(rope_theta), _stream())
    _lib.check(rc, "mb_rope_kv_append")
    return q


def rope3d_kv_append(qkv: torch.Tensor, position_ids3: torch.Tensor, kcache: torch.Tensor, vcache: torch.Tensor,
                     B: int, S: int, H: int, t: int, rope_theta: float, mrope_section=(16, 24, 24),
                     t_dev: torch.Tensor | None = None) -> torch.Tensor:
    """The 3-D M-RoPE variant of rope_kv_append: position_ids3 int32 [3, B*S] (temporal, height, width)."""
    _check_bf16(qkv, kcache, vcache)
    if position_ids3.dtype != torch.int32 or position_ids3.shape != (3, B * S) or not position_ids3.is_contiguous():
        raise ValueError(f"position_ids3 must be a contiguous int32 [3, {B * S}] tensor")
    lib = _lib.load()
    _, Hkv, Tmax, hd = kcache.shape
    s0, s1, s2 = (int(v) for v in mrope_section)
    q = torch.empty((B * S, H * hd), dtype=BF16, device=qkv.device)
    rc = lib.mb_rope3d_kv_append(qkv.data_ptr(), _iptr(position_ids3), q.data_ptr(), kcache.data_ptr(),
                                 vcache.data_ptr(), B, S, H, Hkv, hd, Tmax, _iptr(t_dev), int(t), float(rope_theta),
                                 s0, s1, s2, _stream())
    _lib.check(rc, "mb_rope3d_kv_append")
    return q


def attn_decode_gqa(q: torch.Tensor, kcache: torch.Tensor, vcache: torch.Tensor, key_mask: torch.Tensor | None,
                    H: int, T: int, t_dev: torch.Tensor | None = None) -> torch.Tensor:
    """q [B, H*128] against cache slots 0..T-1 (key_mask int32 [B, >=T], 0 = skip).  Contexts longer than a few
    hundred keys are split over several CTAs per head (flash-decoding) so that B * H = 16..48 heads fill the GPU."""
    _check_bf16(q, kcache, vcache)
    lib = _lib.load()
    B, Hkv, Tmax, hd = kcache.shape
    B = q.shape[0]
    out = torch.empty((B, H * hd), dtype=BF16, device=q.device)
    t_bound = Tmax if t_dev is not None else int(T)   # static upper bound of the context length of this call
    n_splits = max(1, min((t_bound + 63) // 64, max(1, 592 // (B * H))))
    ws = torch.empty((B * H * n_splits * 130,), dtype=torch.float32, device=q.device) if n_splits > 1 else None
    rc = lib.mb_attn_decode_gqa(q.data_ptr(), kcache.data_ptr(), vcache.data_ptr(), _iptr(key_mask),
                                key_mask.stride(0) if key_mask is not None else 0, out.data_ptr(), B, H, Hkv, hd, Tmax,
                                _iptr(t_dev), int(T), hd ** -0.5, _ptr(ws), n_splits, _stream())
    _lib.check(rc, "mb_attn_decode_gqa")
    return out


def attn_prefill_gqa(q: torch.Tensor, kcache: torch.Tensor, vcache: torch.Tensor, B: int, S: int, H: int,
                     t0: int = 0) -> torch.Tensor:
    """Causal GQA attention of S new tokens, already appended at cache slots t0..t0+S-1, over keys 0..t0+S-1 (bottom-right
    aligned causal mask: a later round's prompt sees the whole earlier context) for head_dim 64/128: q [B*S, H*hd]."""
    _check_bf16(q, kcache, vcache)
    lib = _lib.load()
    _, Hkv, Tmax, hd = kcache.shape
    out = torch.empty((B * S, H * hd), dtype=BF16, device=q.device)
    rc = lib.mb_attn_fwd(q.data_ptr(), S * H * hd, H * hd, hd, kcache.data_ptr(), Hkv * Tmax * hd, hd, Tmax * hd,
                         vcache.data_ptr(), Hkv * Tmax * hd, hd, Tmax * hd, out.data_ptr(), S * H * hd, H * hd, hd,
                         B, S, t0 + S, H, Hkv, hd, hd ** -0.5, 1, _stream())
    _lib.check(rc, "mb_attn_fwd")
    return out


def router_topk(logits: torch.Tensor, k: int, renorm: bool, logits_img: torch.Tensor | None = None,
                image_mask: torch.Tensor | None = None) -> tuple[torch.Tensor, torch.Tensor]:
    _check_bf16(logits, logits_img)
    lib = _lib.load()
    T, E = logits.shape
    idx = torch.empty((T, k), dtype=torch.int32, device=logits.device)
    w = torch.empty((T, k), dtype=torch.float32, device=logits.device)
    if image_mask is not None and image_mask.dtype != torch.uint8:
        raise TypeError("image_mask must be uint8")
    rc = lib.mb_router_topk(logits.contiguous().data_ptr(), _ptr(logits_img), _ptr(image_mask), idx.data_ptr(),
                            w.data_ptr(), T, E, k, int(renorm), _stream())
    _lib.check(rc, "mb_router_topk")
    return idx, w


# token count above which the routed experts run as ONE grouped tcgen05 GEMM per projection instead of the
# weight-streaming kernel (which re-reads an expert's 17 MB once per 8 rows); MB_MOE_GROUPED=0/1 forces either path
MOE_GROUPED_MIN_PAIRS_PER_EXPERT = 16


def _moe_use_grouped(T: int, k: int, E: int) -> bool:
    force = os.environ.get("MB_MOE_GROUPED")
    if force is not None:
        return force == "1"
    return T * k >= MOE_GROUPED_MIN_PAIRS_PER_EXPERT * E


def moe_plan(idx: torch.Tensor, E: int, e_begin: int = 0, div: int = 1, granule: int = 128, want_counts: bool = False):
    """Routing plan (mb_moe_plan): idx int32 [T, k] (global expert ids) bucketed by (idx - e_begin) // div into E
    buckets, each padded to `granule` rows -> (pair_row [T*k], row_token [max_rows], tile_expert [max_rows/128] or
    None, meta [2] = {128-row tiles, rows}, max_rows[, counts [E]])."""
    lib = _lib.load()
    T, k = idx.shape
    dev = idx.device
    max_rows = T * k + (granule - 1) * E
    max_rows = ((max_rows + granule - 1) // granule) * granule
    pair_row = torch.empty((T * k,), dtype=torch.int32, device=dev)
    row_token = torch.empty((max_rows,), dtype=torch.int32, device=dev)
    tile_expert = torch.zeros((max_rows // 128,), dtype=torch.int32, device=dev) if granule == 128 else None
    meta = torch.empty((2,), dtype=torch.int32, device=dev)
    counts = torch.empty((E,), dtype=torch.int32, device=dev) if want_counts else None
    _lib.check(lib.mb_moe_plan(idx.data_ptr(), pair_row.data_ptr(), row_token.data_ptr(), _ptr(tile_expert),
                               meta.data_ptr(), _ptr(counts), T, k, E, int(e_begin), int(div), int(granule), max_rows,
                               _stream()), "mb_moe_plan")
    if want_counts:
        return pair_row, row_token, tile_expert, meta, max_rows, counts
    return pair_row, row_token, tile_expert, meta, max_rows


def gather_rows(x: torch.Tensor, row_index: torch.Tensor, n_rows: int, meta: torch.Tensor | None = None) -> torch.Tensor:
    """out[r] = x[row_index[r]] (zero row where the index is negative) for r < n_rows (or < meta[1], read on the device)."""
    _check_bf16(x)
    lib = _lib.load()
    D = x.shape[-1]
    out = torch.empty((n_rows, D), dtype=BF16, device=x.device)
    _lib.check(lib.mb_moe_gather_rows(x.data_ptr(), row_index.data_ptr(), _ptr(meta), out.data_ptr(), n_rows, D,
                                      _stream()), "mb_moe_gather_rows")
    return out


def moe_grouped_ffn(xg: torch.Tensor, Wgu: torch.Tensor, Wd: torch.Tensor, tile_expert: torch.Tensor,
                    meta: torch.Tensor) -> torch.Tensor:
    """Expert FFN over expert-sorted, 128-row-padded rows xg [max_rows, D]: grouped gate/up GEMM with the SwiGLU
    epilogue, then the grouped down GEMM (both tcgen05 / TMA, one launch each) -> [max_rows, D]."""
    _check_bf16(xg, Wgu, Wd)
    lib = _lib.load()
    max_rows, D = xg.shape
    E, I2, _ = Wgu.shape
    I = I2 // 2
    s = _stream()
    hid = torch.empty((max_rows, I), dtype=BF16, device=xg.device)
    out = torch.empty((max_rows, D), dtype=BF16, device=xg.device)
    _lib.check(lib.mb_moe_grouped_gemm(xg.data_ptr(), Wgu.data_ptr(), hid.data_ptr(), tile_expert.data_ptr(),
                                       meta.data_ptr(), max_rows, I2, D, E, 1, s), "mb_moe_grouped_gemm")
    _lib.check(lib.mb_moe_grouped_gemm(hid.data_ptr(), Wd.data_ptr(), out.data_ptr(), tile_expert.data_ptr(),
                                       meta.data_ptr(), max_rows, D, I, E, 0, s), "mb_moe_grouped_gemm")
    return out


def moe_experts(x: torch.Tensor, idx: torch.Tensor, w: torch.Tensor, Wgu: torch.Tensor, Wd: torch.Tensor,
                shared: torch.Tensor | None, residual: torch.Tensor | None, e_begin: int = 0,
                ep_group=None, ep_peer=None) -> torch.Tensor:
    """moe_infer + shared-expert add + residual: x [T, D]; Wgu [E_local, 2I, D]; Wd [E_local, D, I]; idx int32 [T, k]
    (GLOBAL expert ids); w fp32.  Small token counts stream each hit expert's weights once per 8 rows
    (mb_moe_gate_up / mb_moe_down); prefill-sized inputs run as grouped tcgen05 GEMMs (mb_moe_plan /
    mb_moe_grouped_gemm).  With `ep_group` the slabs hold only experts [e_begin, e_begin + E_local): each rank
    computes the fp32 weighted sum of its experts, the partials are all-reduced over NCCL (torch.distributed is the
    plumbing; the tokens of this path are replicated on every rank, so there is no dispatch exchange), and the
    reference's rounding chain is applied after the reduction."""
    _check_bf16(x, Wgu, Wd, shared, residual)
    lib = _lib.load()
    T, D = x.shape
    E, I2, _ = Wgu.shape
    I = I2 // 2
    k = idx.shape[1]
    dev = x.device
    ep = ep_group is not None
    y = torch.empty((T, D), dtype=BF16, device=dev)
    s = _stream()
    x = x.contiguous()
    if _moe_use_grouped(T, k, E):
        pair_row, row_token, tile_expert, meta, max_rows = moe_plan(idx, E, e_begin)
        xg = gather_rows(x, row_token, max_rows, meta)
        out_pairs = moe_grouped_ffn(xg, Wgu, Wd, tile_expert, meta)
        pr = pair_row.data_ptr()
    elif not ep:
        # decode-sized, all experts local: the whole-stage C driver (sort -> gate/up -> down -> combine, one call)
        import ctypes

        nbytes = ctypes.c_int64(0)
        _lib.check(lib.mb_moe_ffn_workspace_bytes(T, k, E, D, I, ctypes.byref(nbytes)), "mb_moe_ffn_workspace_bytes")
        ws = torch.empty((nbytes.value,), dtype=torch.uint8, device=dev)
        _lib.check(lib.mb_moe_ffn(x.data_ptr(), idx.data_ptr(), w.data_ptr(), Wgu.data_ptr(), Wd.data_ptr(), _ptr(shared),
                                  _ptr(residual), y.data_ptr(), ws.data_ptr(), ws.numel(), T, k, E, int(e_begin), E, D, I,
                                  s), "mb_moe_ffn")
        _lib.count_replay(3)  # (four kernels behind one C call)
        return y
    else:
        offs = torch.empty((E + 1,), dtype=torch.int32, device=dev)
        sorted_pair = torch.empty((T * k,), dtype=torch.int32, device=dev)
        hid = torch.empty((T * k, I), dtype=BF16, device=dev)
        out_pairs = torch.zeros((T * k, D), dtype=BF16, device=dev)
        _lib.check(lib.mb_moe_sort(idx.data_ptr(), offs.data_ptr(), sorted_pair.data_ptr(), T, k, E, int(e_begin), s),
                   "mb_moe_sort")
        # (with ep_group the slabs hold E of E * world experts; the pairs spread over all of them)
        n_all = E
        if ep:
            import torch.distributed as dist

            n_all = E * dist.get_world_size(ep_group)
        mean_pairs = -(-T * k // n_all)
        _lib.check(lib.mb_moe_gate_up(x.data_ptr(), Wgu.data_ptr(), offs.data_ptr(), sorted_pair.data_ptr(),
                                      hid.data_ptr(), T, k, E, D, I, mean_pairs, s), "mb_moe_gate_up")
        _lib.check(lib.mb_moe_down(hid.data_ptr(), Wd.data_ptr(), offs.data_ptr(), sorted_pair.data_ptr(),
                                   out_pairs.data_ptr(), T, k, E, D, I, mean_pairs, s), "mb_moe_down")
        pr = None
    if not ep:
        _lib.check(lib.mb_moe_combine(out_pairs.data_ptr(), w.data_ptr(), _ptr(shared), _ptr(residual), y.data_ptr(),
                                      None, pr, T, k, D, s), "mb_moe_combine")
        return y
    if ep_peer is not None and T <= ep_peer.T_MAX:
        # fused exchange over NVLink peer memory: partial sums are stored straight into every peer's area, then every
        # rank adds the G partials in rank order (no NCCL call, no host sync)
        if D != ep_peer.hidden_size:
            raise ValueError("peer exchange area was sized for another hidden size")
        _lib.check(lib.mb_moe_combine_push(out_pairs.data_ptr(), w.data_ptr(), pr, ep_peer.peers_dev, ep_peer.rank,
                                           ep_peer.size, T, ep_peer.T_MAX, k, D, s), "mb_moe_combine_push")
        _lib.check(lib.mb_moe_reduce_finalize(ep_peer.peers_dev, ep_peer.rank, ep_peer.size, T, ep_peer.T_MAX, D,
                                              _ptr(shared), _ptr(residual), y.data_ptr(), ep_peer.fin_done.data_ptr(),
                                              s), "mb_moe_reduce_finalize")
        return y
    import torch.distributed as dist

    part = torch.empty((T, D), dtype=torch.float32, device=dev)
    _lib.check(lib.mb_moe_combine(out_pairs.data_ptr(), w.data_ptr(), None, None, y.data_ptr(), part.data_ptr(), pr, T,
                                  k, D, s), "mb_moe_combine")
    dist.all_reduce(part, op=dist.ReduceOp.SUM, group=ep_group)
    _lib.check(lib.mb_moe_finalize(part.data_ptr(), _ptr(shared), _ptr(residual), y.data_ptr(), T, D, s),
               "mb_moe_finalize")
    return y


def moe_local_experts(x: torch.Tensor, idx: torch.Tensor, Wgu: torch.Tensor, Wd: torch.Tensor, e_begin: int,
                      n_experts_total: int | None = None):
    """The routed-expert FFNs of the LOCAL experts [e_begin, e_begin + E_local) on rows x [T, D] with GLOBAL expert ids
    idx [T, k]: returns (out_pairs, pair_row) — out_pairs rows in (token, slot) order with pair_row None (streaming
    kernels; rows of non-local pairs are never written), or in the grouped layout addressed through pair_row (negative =
    not local).  The building block of the expert-parallel paths; the weighted combine is the caller's."""
    _check_bf16(x, Wgu, Wd)
    lib = _lib.load()
    T, D = x.shape
    E, I2, _ = Wgu.shape
    I = I2 // 2
    k = idx.shape[1]
    dev = x.device
    s = _stream()
    # the pairs spread over ALL experts, E of which are local: pairs per local expert ~ T k / n_experts_total
    if _moe_use_grouped(T, k, n_experts_total or E):
        pair_row, row_token, tile_expert, meta, max_rows = moe_plan(idx, E, e_begin)
        xg = gather_rows(x, row_token, max_rows, meta)
        return moe_grouped_ffn(xg, Wgu, Wd, tile_expert, meta), pair_row
    offs = torch.empty((E + 1,), dtype=torch.int32, device=dev)
    sorted_pair = torch.empty((T * k,), dtype=torch.int32, device=dev)
    hid = torch.empty((T * k, I), dtype=BF16, device=dev)
    out_pairs = torch.empty((T * k, D), dtype=BF16, device=dev)
    _lib.check(lib.mb_moe_sort(idx.data_ptr(), offs.data_ptr(), sorted_pair.data_ptr(), T, k, E, int(e_begin), s),
               "mb_moe_sort")
    mean_pairs = -(-T * k // (n_experts_total or E))
    _lib.check(lib.mb_moe_gate_up(x.data_ptr(), Wgu.data_ptr(), offs.data_ptr(), sorted_pair.data_ptr(),
                                  hid.data_ptr(), T, k, E, D, I, mean_pairs, s), "mb_moe_gate_up")
    _lib.check(lib.mb_moe_down(hid.data_ptr(), Wd.data_ptr(), offs.data_ptr(), sorted_pair.data_ptr(),
                               out_pairs.data_ptr(), T, k, E, D, I, mean_pairs, s), "mb_moe_down")
    return out_pairs, None


# ---- expert parallelism over peer memory, phase by phase (ep.PeerDispatch; csrc/ep.cu).  A real rank runs the four
# phases back to back; the single-device test drives G virtual ranks phase by phase.
def ep_dispatch(pd, x: torch.Tensor, idx: torch.Tensor, w: torch.Tensor, rank: int | None = None) -> None:
    _check_bf16(x)
    T, D = x.shape
    if T > pd.t_max or D != pd.hidden_size or idx.shape != (T, pd.top_k) or idx.dtype != torch.int32:
        raise ValueError(f"expert-parallel dispatch: rows {tuple(x.shape)} / idx {tuple(idx.shape)} do not fit the "
                         f"exchange area (t_max {pd.t_max}, D {pd.hidden_size}, k {pd.top_k})")
    r = pd.rank if rank is None else rank
    _lib.check(_lib.load().mb_ep_dispatch_push(x.contiguous().data_ptr(), idx.contiguous().data_ptr(),
                                               w.contiguous().data_ptr(), pd.peers_dev, r, pd.size, T, pd.t_max, D,
                                               pd.top_k, _stream()), "mb_ep_dispatch_push")


def ep_compute(pd, T: int, Wgu: torch.Tensor, Wd: torch.Tensor, e_begin: int, n_experts_total: int,
               rank: int | None = None):
    r = pd.rank if rank is None else rank
    lib = _lib.load()
    x_all, idx_all, _ = pd.gathered(T, r)
    Ta, k = x_all.shape[0], pd.top_k
    E, I2, D = Wgu.shape
    if _moe_use_grouped(Ta, k, n_experts_total) or (2 * E + Ta * k) * 4 > 48 * 1024:
        # prefill-sized: wait, then the tile plan + grouped tcgen05 GEMMs (or the plain sort when the pairs are too many
        # for the fused kernel's shared memory)
        _lib.check(lib.mb_ep_dispatch_wait(pd.peers_dev, r, pd.size, pd.t_max, pd.hidden_size, k, _stream()),
                   "mb_ep_dispatch_wait")
        return moe_local_experts(x_all, idx_all, Wgu, Wd, e_begin, n_experts_total)
    # decode-sized: the wait rides in the sort kernel of the streaming path
    dev, s = x_all.device, _stream()
    offs = torch.empty((E + 1,), dtype=torch.int32, device=dev)
    sorted_pair = torch.empty((Ta * k,), dtype=torch.int32, device=dev)
    hid = torch.empty((Ta * k, I2 // 2), dtype=BF16, device=dev)
    out_pairs = torch.empty((Ta * k, D), dtype=BF16, device=dev)
    _lib.check(lib.mb_ep_wait_sort(pd.peers_dev, r, pd.size, T, pd.t_max, pd.hidden_size, k, offs.data_ptr(),
                                   sorted_pair.data_ptr(), E, int(e_begin), s), "mb_ep_wait_sort")
    mean_pairs = -(-Ta * k // n_experts_total)
    _lib.check(lib.mb_moe_gate_up(x_all.data_ptr(), Wgu.data_ptr(), offs.data_ptr(), sorted_pair.data_ptr(),
                                  hid.data_ptr(), Ta, k, E, D, I2 // 2, mean_pairs, s), "mb_moe_gate_up")
    _lib.check(lib.mb_moe_down(hid.data_ptr(), Wd.data_ptr(), offs.data_ptr(), sorted_pair.data_ptr(),
                               out_pairs.data_ptr(), Ta, k, E, D, I2 // 2, mean_pairs, s), "mb_moe_down")
    return out_pairs, None


def ep_combine(pd, T: int, out_pairs: torch.Tensor, pair_row: torch.Tensor | None, e_begin: int, e_local: int,
               rank: int | None = None) -> None:
    r = pd.rank if rank is None else rank
    _lib.check(_lib.load().mb_ep_combine_push(out_pairs.data_ptr(), _ptr(pair_row), pd.peers_dev, r, pd.size, T,
                                              pd.t_max, pd.hidden_size, pd.top_k, int(e_begin), int(e_local),
                                              _stream()), "mb_ep_combine_push")


def ep_finalize(pd, T: int, shared: torch.Tensor | None, residual: torch.Tensor | None,
                rank: int | None = None) -> torch.Tensor:
    _check_bf16(shared, residual)
    r = pd.rank if rank is None else rank
    y = torch.empty((T, pd.hidden_size), dtype=BF16, device=pd.area(r).device)
    _lib.check(_lib.load().mb_ep_reduce_finalize(pd.peers_dev, r, pd.size, T, pd.t_max, pd.hidden_size, pd.top_k,
                                                 _ptr(shared), _ptr(residual), y.data_ptr(), _stream()),
               "mb_ep_reduce_finalize")
    return y


def moe_combine(out_pairs: torch.Tensor, w: torch.Tensor, shared: torch.Tensor | None, residual: torch.Tensor | None,
                y: torch.Tensor, pair_row: torch.Tensor | None = None) -> torch.Tensor:
    """y[t] = bf16(bf16(bf16(sum_j w[t, j] * out_pairs[row(t, j)]) + shared[t]) + residual[t]) written into `y` [T, D];
    row(t, j) = pair_row[t*k + j] (or t*k + j without a plan)."""
    _check_bf16(out_pairs, shared, residual, y)
    if not y.is_contiguous():
        raise ValueError("moe_combine writes a contiguous [T, D] view")
    lib = _lib.load()
    T, k = w.shape
    _lib.check(lib.mb_moe_combine(out_pairs.data_ptr(), w.data_ptr(), _ptr(shared), _ptr(residual), y.data_ptr(), None,
                                  _ptr(pair_row), T, k, y.shape[1], _stream()), "mb_moe_combine")
    return y


def argmax_rows(logits: torch.Tensor) -> torch.Tensor:
    """fp32 [rows, V] -> int32 [rows] (first index of the maximum)."""
    if logits.dtype != torch.float32 or not logits.is_cuda or not logits.is_contiguous():
        raise TypeError("argmax_rows expects a contiguous CUDA fp32 tensor")
    lib = _lib.load()
    rows, V = logits.shape
    out = torch.empty((rows,), dtype=torch.int32, device=logits.device)
    n_chunks = max(1, min(128, V // 1024))
    wv = torch.empty((rows * n_chunks,), dtype=torch.float32, device=logits.device) if n_chunks > 1 else None
    wi = torch.empty((rows * n_chunks,), dtype=torch.int32, device=logits.device) if n_chunks > 1 else None
    _lib.check(lib.mb_argmax_f32(logits.data_ptr(), out.data_ptr(), rows, V, _ptr(wv), _ptr(wi), n_chunks, _stream()),
               "mb_argmax_f32")
    return out


# ---------------------------------------------------------------------------------------------------------------
# image pre- / post-processing
# ---------------------------------------------------------------------------------------------------------------
def resized_output_size(h: int, w: int, size) -> tuple[int, int]:
    """torchvision's Resize rule (`_compute_resized_output_size`, no max_size): an int (or 1-tuple) sets the SHORT edge
    and the long edge becomes int(size * long / short); a pair is (h, w) verbatim."""
    if isinstance(size, (tuple, list)):
        if len(size) == 2:
            return int(size[0]), int(size[1])
        if len(size) != 1:
            raise ValueError(f"size must be an int or a 1- or 2-element sequence, not {size!r}")
        size = size[0]
    size = int(size)
    short, long = (w, h) if w <= h else (h, w)
    new_short, new_long = size, int(size * long / short)
    new_w, new_h = (new_short, new_long) if w <= h else (new_long, new_short)
    return new_h, new_w


def center_crop_window(h: int, w: int, crop_h: int, crop_w: int) -> tuple[int, int]:
    """torchvision `center_crop` offsets (top, left) = int(round((full - crop) / 2.0)) — Python's round, half to even.
    Crops larger than the image (torchvision zero-pads) do not occur after Resize(short edge) and are refused."""
    if crop_h > h or crop_w > w:
        raise ValueError(f"centre crop {crop_h}x{crop_w} larger than the resized image {h}x{w}")
    return int(round((h - crop_h) / 2.0)), int(round((w - crop_w) / 2.0))


def image_preprocess(images: torch.Tensor, size, crop: int | tuple[int, int] | None = None,
                     mean=(0.5, 0.5, 0.5), std=(0.5, 0.5, 0.5), out_dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """Resize(size, bicubic, Pillow's antialiased 8-bit resample) [-> CenterCrop(crop)] -> ToTensor -> Normalize on the
    device: images uint8 [N, H, W, 3] (or [H, W, 3]) RGB on CUDA -> [N, 3, h, w] `out_dtype` (fp32: identical to the
    torchvision pipeline on PIL images; bf16: its round-to-nearest; uint8: the resized + cropped image itself,
    [N, h, w, 3], without ToTensor / Normalize)."""
    if images.dtype != torch.uint8 or not images.is_cuda:
        raise TypeError(f"expected a CUDA uint8 tensor, got {images.dtype} on {images.device}")
    if images.dim() == 3:
        images = images.unsqueeze(0)
    if images.dim() != 4 or images.shape[-1] != 3:
        raise ValueError(f"expected [N, H, W, 3] RGB images, got {tuple(images.shape)}")
    if out_dtype not in (torch.float32, BF16, torch.uint8):
        raise TypeError("out_dtype must be float32, bfloat16 or uint8")
    images = images.contiguous()
    n, h, w, _ = images.shape
    rh, rw = resized_output_size(h, w, size)
    if crop is None:
        top, left, oh, ow = 0, 0, rh, rw
    else:
        oh, ow = (int(crop), int(crop)) if isinstance(crop, int) else (int(crop[0]), int(crop[1]))
        top, left = center_crop_window(rh, rw, oh, ow)
    import ctypes

    lib = _lib.load()
    nbytes = ctypes.c_int64(0)
    _lib.check(lib.mb_image_preprocess_workspace_bytes(n, h, w, rh, rw, top, left, oh, ow, ctypes.byref(nbytes)),
               "mb_image_preprocess_workspace_bytes")
    ws = torch.empty((max(16, nbytes.value),), dtype=torch.uint8, device=images.device)
    kind = {BF16: 0, torch.float32: 1, torch.uint8: 2}[out_dtype]
    shape = (n, oh, ow, 3) if kind == 2 else (n, 3, oh, ow)
    out = torch.empty(shape, dtype=out_dtype, device=images.device)
    _lib.check(lib.mb_image_preprocess_u8(images.data_ptr(), n, h, w, rh, rw, top, left, oh, ow, float(mean[0]),
                                          float(mean[1]), float(mean[2]), float(std[0]), float(std[1]), float(std[2]),
                                          out.data_ptr(), kind, ws.data_ptr(), ws.numel(), _stream()),
               "mb_image_preprocess_u8")
    return out


def image_resize_u8(images: torch.Tensor, size, crop: int | tuple[int, int] | None = None) -> torch.Tensor:
    """Pillow's `Image.resize(BICUBIC)` (+ torchvision's centre crop) on the device, u8 in / u8 out: uint8 [N, H, W, 3]
    -> uint8 [N, h, w, 3], bit-identical with the library (the resize `fetch_image` applies before the processors,
    mingunivision/bailingmm_utils.py:162, is this call with size = (resized_height, resized_width))."""
    return image_preprocess(images, size, crop, out_dtype=torch.uint8)


def image_postprocess(img: torch.Tensor, mean=(0.5, 0.5, 0.5), std=(0.5, 0.5, 0.5)) -> torch.Tensor:
    """`tensor_to_pil` on the device: [N, 3, H, W] (or [3, H, W]) fp32 / bf16 in [-1, 1] -> uint8 [N, H, W, 3] =
    trunc((x * std + mean) * 255)."""
    if img.dtype not in (torch.float32, BF16) or not img.is_cuda:
        raise TypeError(f"expected a CUDA float32 / bfloat16 tensor, got {img.dtype} on {img.device}")
    if img.dim() == 3:
        img = img.unsqueeze(0)
    if img.dim() != 4 or img.shape[1] != 3:
        raise ValueError(f"expected [N, 3, H, W], got {tuple(img.shape)}")
    img = img.contiguous()
    n, _, h, w = img.shape
    out = torch.empty((n, h, w, 3), dtype=torch.uint8, device=img.device)
    lib = _lib.load()
    _lib.check(lib.mb_image_postprocess_u8(img.data_ptr(), int(img.dtype == torch.float32), n, h, w, float(mean[0]),
                                           float(mean[1]), float(mean[2]), float(std[0]), float(std[1]), float(std[2]),
                                           out.data_ptr(), _stream()), "mb_image_postprocess_u8")
    return out


def unpatchify_to_u8(x: torch.Tensor, g: int, p: int, mean=(0.5, 0.5, 0.5), std=(0.5, 0.5, 0.5)) -> torch.Tensor:
    """Head rows [B, g*g, p*p*3] bf16 -> uint8 [B, g*p, g*p, 3]: unpatchify + clamp(-1, 1) + `tensor_to_pil`'s
    conversion in one pass (== image_postprocess(unpatchify_clamp(x)))."""
    _check_bf16(x)
    lib = _lib.load()
    x = x.contiguous()
    B = x.shape[0]
    if x.shape[1] != g * g or x.shape[2] != p * p * 3:
        raise ValueError(f"expected [B, {g * g}, {p * p * 3}], got {tuple(x.shape)}")
    out = torch.empty((B, g * p, g * p, 3), dtype=torch.uint8, device=x.device)
    _lib.check(lib.mb_unpatchify_to_u8(x.data_ptr(), out.data_ptr(), B, g, p, float(mean[0]), float(mean[1]),
                                       float(mean[2]), float(std[0]), float(std[1]), float(std[2]), _stream()),
               "mb_unpatchify_to_u8")
    return out
