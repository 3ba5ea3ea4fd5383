"""Operator-level numerics: every CUDA kernel behind the C ABI against a plain PyTorch fp32 reference of the same
op on the same inputs (floating-point kernels; tolerances are the bf16 output rounding, stated per test)."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

BF16 = torch.bfloat16


def _rand(shape, dev, scale=1.0, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(dev).to(BF16)


def _rel_err(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


GEMM_SHAPES = [
    # (M, N, K)                      covers: single tile, M/N/K tails, both tile widths, long K, many tiles
    (128, 256, 64), (128, 128, 128), (1, 32, 64), (65, 768, 768), (130, 2304, 768), (257, 1024, 2816),
    (4160, 768, 2048), (1000, 96, 264), (300, 1000, 72), (16384, 1024, 1024), (2048, 3072, 1024), (512, 4096, 1024),
]


@pytest.fixture(params=["auto", "pair256", "pair128", "single256", "single128", "pair256-direct", "single128-direct"])
def tile(request):
    """Pins the GEMM tile shape (CTA pair = tcgen05 cta_group::2) and the epilogue flavour (staged TMA store vs direct
    register stores) so every kernel variant sees every shape."""
    from ming_univision_b200 import _lib

    cg, bn = {"auto": (0, 0), "pair256": (2, 256), "pair128": (2, 128), "single256": (1, 256),
              "single128": (1, 128), "pair256-direct": (2 + 16, 256), "single128-direct": (1 + 16, 128)}[request.param]
    _lib.check(_lib.load().mb_gemm_force_tile(cg, bn), "mb_gemm_force_tile")
    yield request.param
    _lib.load().mb_gemm_force_tile(0, 0)


@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
@pytest.mark.parametrize("epi", ["bias", "gelu", "residual"])
def test_gemm_epilogues(cuda_device, tile, M, N, K, epi):
    from ming_univision_b200 import ops

    if epi != "bias" and M * N * K > 4e9:
        pytest.skip("large shape covered by the bias epilogue")
    x = _rand((M, K), cuda_device, 1.0, 1)
    w = _rand((N, K), cuda_device, 1.0 / math.sqrt(K), 2)
    b = _rand((N,), cuda_device, 0.5, 3)
    ref = x.float() @ w.float().t() + b.float()
    pre = ref
    if epi == "bias":
        out = ops.linear(x, w, b)
    elif epi == "gelu":
        out = ops.linear(x, w, b, epi=ops.EPI_GELU)
        ref = F.gelu(ref.to(BF16).float())
    else:
        r = _rand((M, N), cuda_device, 1.0, 4)
        out = ops.linear(x, w, b, epi=ops.EPI_RESIDUAL, residual=r)
        ref = ref.to(BF16).float() + r.float()
    torch.cuda.synchronize()
    assert out.shape == (M, N)
    # bf16 rounding of the output (2^-8 relative) plus one bf16 ulp of the pre-activation / pre-residual value, whose
    # rounding to bf16 (mirroring the reference's dtype flow) can flip when fp32 summation order differs
    err = (out.float() - ref).abs()
    tol = 2.0 ** -7 * ref.abs() + 2.0 ** -7 * pre.abs() + 1e-3
    assert (err <= tol).all(), f"max err {err.max().item()} rel {_rel_err(out, ref)}"
    assert _rel_err(out, ref) < 4e-3


@pytest.mark.parametrize("direct", [False, True])
def test_gemm_inplace_residual(cuda_device, direct):
    """out aliases the residual (how the transformer blocks update their residual stream)."""
    from ming_univision_b200 import _lib, ops

    _lib.load().mb_gemm_force_tile(16 if direct else 0, 0)
    M, N, K = 4160, 1024, 1024
    x = _rand((M, K), cuda_device, 1.0, 1)
    w = _rand((N, K), cuda_device, 1.0 / math.sqrt(K), 2)
    b = _rand((N,), cuda_device, 0.5, 3)
    r = _rand((M, N), cuda_device, 1.0, 4)
    ref = (x.float() @ w.float().t() + b.float()).to(BF16).float() + r.float()
    stream = r.clone()
    ops.linear(x, w, b, epi=ops.EPI_RESIDUAL, residual=stream, out=stream)
    _lib.load().mb_gemm_force_tile(0, 0)
    assert _rel_err(stream, ref) < 4e-3
    assert ((stream.float() - ref).abs() <= 2.0 ** -6 * ref.abs() + 2e-2).all()


def test_gemm_no_bias_and_strided(cuda_device):
    from ming_univision_b200 import ops

    x_full = _rand((200, 512), cuda_device, 1.0, 5)
    x = x_full[:, :256]  # row stride 512, K = 256
    w = _rand((384, 256), cuda_device, 1 / 16, 6)
    out = ops.linear(x, w, None)
    ref = x.float() @ w.float().t()
    assert _rel_err(out, ref) < 4e-3


@pytest.mark.parametrize("M,K,H", [(65, 768, 2048), (130, 1024, 2736), (3, 3072, 8192), (4160, 1024, 2736)])
@pytest.mark.parametrize("pair", [True, False])
def test_gemm_swiglu(cuda_device, M, K, H, pair):
    from ming_univision_b200 import _lib, ops

    _lib.load().mb_gemm_force_tile(2 if pair else 1, 256)

    x = _rand((M, K), cuda_device, 1.0, 7)
    w12 = _rand((2 * H, K), cuda_device, 1.0 / math.sqrt(K), 8)
    b12 = _rand((2 * H,), cuda_device, 0.2, 9)
    wp, bp, Hp = ops.pack_swiglu(w12, b12)
    assert Hp % 128 == 0 and wp.shape == (2 * Hp, K)
    out = ops.linear(x, wp, bp, epi=ops.EPI_SWIGLU)
    assert out.shape == (M, Hp)
    x12 = (x.float() @ w12.float().t() + b12.float()).to(BF16).float()
    x1, x2 = x12[:, :H], x12[:, H:]
    ref = F.silu(x1).to(BF16).float() * x2
    torch.cuda.synchronize()
    _lib.load().mb_gemm_force_tile(0, 0)
    assert (out[:, H:].float() == 0).all(), "padded hidden columns must be exactly zero"
    err = (out[:, :H].float() - ref).abs()
    tol = 2.0 ** -6 * ref.abs() + 5e-3
    assert (err <= tol).all(), f"max err {err.max().item()}"
    assert _rel_err(out[:, :H], ref) < 6e-3


def test_gemm_row_remap_and_rowmod_residual(cuda_device):
    """patch-embed use: rows of image b land at b*(n+1)+p and the residual (pos-embed) is indexed by p."""
    from ming_univision_b200 import ops

    B, n, K, N = 5, 64, 3072, 768
    x = _rand((B * n, K), cuda_device, 1.0, 10)
    w = _rand((N, K), cuda_device, 1 / math.sqrt(K), 11)
    b = _rand((N,), cuda_device, 0.1, 12)
    pos = _rand((n, N), cuda_device, 1.0, 13)
    out = torch.full((B, n + 1, N), 7.0, dtype=BF16, device=cuda_device)
    ops.linear(x, w, b, epi=ops.EPI_RESIDUAL, residual=pos, res_row_mod=n, out=out, out_row_group=n, out_row_pad=1)
    ref = (x.float() @ w.float().t() + b.float()).to(BF16).float().view(B, n, N) + pos.float()
    assert (out[:, n].float() == 7.0).all(), "cls rows must be untouched"
    assert _rel_err(out[:, :n], ref) < 4e-3


@pytest.mark.parametrize("rows,dim", [(1, 768), (65, 1024), (4160, 768), (7, 3072), (33, 2048), (10, 64), (9, 136)])
@pytest.mark.parametrize("act", [0, 1])
def test_layernorm(cuda_device, rows, dim, act):
    from ming_univision_b200 import ops

    x = _rand((rows, dim), cuda_device, 2.0, 20) + 0.5
    g = (_rand((dim,), cuda_device, 0.1, 21).float() + 1).to(BF16)
    b = _rand((dim,), cuda_device, 0.1, 22)
    y = ops.layernorm(x, g, b, 1e-6, act)
    ref = pre = F.layer_norm(x.float(), (dim,), g.float(), b.float(), 1e-6)
    if act:
        ref = F.gelu(ref.to(BF16).float())
    err = (y.float() - ref).abs()
    # output rounding + (act only) one bf16 ulp of the pre-activation, whose rounding may flip
    assert (err <= 2.0 ** -7 * ref.abs() + act * 2.0 ** -7 * pre.abs() + 2e-3).all(), f"max err {err.max().item()}"
    y2 = ops.layernorm(x, None, None, 1e-6, 0)
    ref2 = F.layer_norm(x.float(), (dim,), None, None, 1e-6)
    assert ((y2.float() - ref2).abs() <= 2.0 ** -7 * ref2.abs() + 2e-3).all()


@pytest.mark.parametrize("B,S,H", [(1, 65, 12), (2, 257, 16), (3, 64, 16), (2, 256, 16), (1, 1025, 12), (2, 1, 4),
                                   (1, 130, 2)])
@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("backend", [1, 2])
def test_attention_hd64(cuda_device, B, S, H, causal, backend):
    """backend 1: tcgen05 / TMEM kernel (attention_tc.cu), 2: warp-level mma.sync kernel (attention.cu)."""
    from ming_univision_b200 import ops

    qkv = _rand((B, S, 3 * H * 64), cuda_device, 1.0, 30)
    ops.set_attn_backend(backend)
    try:
        out = ops.attention_hd64(qkv, B, S, H, causal)
    finally:
        ops.set_attn_backend(0)
    q, k, v = qkv.float().view(B, S, 3, H, 64).permute(2, 0, 3, 1, 4)
    att = (q @ k.transpose(-1, -2)) * 64 ** -0.5
    if causal:
        att = att.masked_fill(torch.triu(torch.ones(S, S, device=cuda_device, dtype=torch.bool), 1), float("-inf"))
    ref = (att.softmax(-1) @ v).transpose(1, 2).reshape(B, S, H * 64)
    err = (out.float() - ref).abs()
    # P is rounded to bf16 before the PV product (as flash-attn does): 2^-8 relative on O(1) values
    assert err.max().item() < 2e-2, f"max err {err.max().item()}"
    assert _rel_err(out, ref) < 8e-3


@pytest.mark.parametrize("causal", [False, True])
def test_attention_hd64_growing_scores(cuda_device, causal):
    """Scores that keep growing along the key axis: every later key block exceeds the running maximum by far more than
    the tcgen05 kernel's lazy-rescaling bound, so its redo path (true block maximum + rescale of O) is exercised."""
    from ming_univision_b200 import ops

    B, S, H = 2, 500, 3
    qkv = _rand((B, S, 3, H, 64), cuda_device, 1.0, 33).float()
    ramp = torch.linspace(0.2, 9.0, S, device=cuda_device).view(1, S, 1, 1)
    qkv[:, :, 1] = qkv[:, :, 1].abs() * ramp          # keys: positive, growing norm
    qkv[:, :, 0] = qkv[:, :, 0].abs() * 2.0           # queries: positive -> scores grow with the key index
    qkv = qkv.to(BF16).view(B, S, 3 * H * 64)
    outs = {}
    for backend in (1, 2):
        ops.set_attn_backend(backend)
        try:
            outs[backend] = ops.attention_hd64(qkv, B, S, H, causal)
        finally:
            ops.set_attn_backend(0)
    q, k, v = qkv.float().view(B, S, 3, H, 64).permute(2, 0, 3, 1, 4)
    att = (q @ k.transpose(-1, -2)) * 64 ** -0.5
    assert (att[..., -1] - att[..., 0]).min().item() > 50  # beyond the bound of 8 / (scale log2 e) = 44
    if causal:
        att = att.masked_fill(torch.triu(torch.ones(S, S, device=cuda_device, dtype=torch.bool), 1), float("-inf"))
    ref = (att.softmax(-1) @ v).transpose(1, 2).reshape(B, S, H * 64)
    for backend in (1, 2):
        assert torch.isfinite(outs[backend].float()).all()
        assert (outs[backend].float() - ref).abs().max().item() < 3e-2, backend
        assert _rel_err(outs[backend], ref) < 8e-3, backend


@pytest.mark.parametrize("B,S,H,Hkv,hd", [(1, 300, 16, 4, 128), (2, 129, 8, 8, 128), (1, 1552, 16, 4, 128),
                                          (2, 77, 4, 2, 64), (1, 40, 4, 1, 128)])
@pytest.mark.parametrize("backend", [1, 2])
def test_attention_prefill_gqa(cuda_device, B, S, H, Hkv, hd, backend):
    """Causal GQA prefill attention reading K / V in place from the [B, Hkv, Tmax, hd] cache (mb_attn_fwd)."""
    from ming_univision_b200 import ops

    Tmax = S + 19
    q = _rand((B * S, H * hd), cuda_device, 1.0, 40)
    kc = _rand((B, Hkv, Tmax, hd), cuda_device, 1.0, 41)
    vc = _rand((B, Hkv, Tmax, hd), cuda_device, 1.0, 42)
    ops.set_attn_backend(backend)
    try:
        out = ops.attn_prefill_gqa(q, kc, vc, B, S, H)
    finally:
        ops.set_attn_backend(0)
    qf = q.float().view(B, S, H, hd).permute(0, 2, 1, 3)
    kf = kc[:, :, :S].float().repeat_interleave(H // Hkv, dim=1)
    vf = vc[:, :, :S].float().repeat_interleave(H // Hkv, dim=1)
    att = (qf @ kf.transpose(-1, -2)) * hd ** -0.5
    att = att.masked_fill(torch.triu(torch.ones(S, S, device=cuda_device, dtype=torch.bool), 1), float("-inf"))
    ref = (att.softmax(-1) @ vf).permute(0, 2, 1, 3).reshape(B * S, H * hd)
    assert (out.float() - ref).abs().max().item() < 2e-2
    assert _rel_err(out, ref) < 8e-3


def test_attention_decode_matches_full_causal(cuda_device):
    from ming_univision_b200 import ops

    B, H, T = 3, 16, 40
    qkv = _rand((B, T, 3 * H * 64), cuda_device, 1.0, 31)
    full = ops.attention_hd64(qkv, B, T, H, True)
    kc = torch.zeros((B, H, 64, 64), dtype=BF16, device=cuda_device)
    vc = torch.zeros_like(kc)
    for t in range(T):
        o = ops.attention_hd64_decode(qkv[:, t].contiguous(), kc, vc, t)
        assert (o.float() - full[:, t].float()).abs().max().item() < 2e-2
    k_ref = qkv.view(B, T, 3, H, 64)[:, :, 1].permute(0, 2, 1, 3)
    assert torch.equal(kc[:, :, :T], k_ref)


def test_mingtok_data_movement(cuda_device):
    from einops import rearrange

    from ming_univision_b200 import ops

    B, C, P, g = 3, 3, 32, 4
    img = torch.randn((B, C, g * P, g * P), device=cuda_device)
    rows = ops.patchify(img, P)
    w = torch.randn((16, C, P, P), device=cuda_device)
    # fp64 on the CPU: cuDNN's fp32 conv may run in TF32, which is not a reference
    ref = F.conv2d(img.to(BF16).double().cpu(), w.double().cpu(), stride=P).flatten(2).transpose(1, 2)  # [B, n, 16]
    got = rows.double().cpu().view(B, g * g, -1) @ w.double().cpu().flatten(1).t()
    assert torch.allclose(got, ref, atol=1e-9, rtol=1e-9)
    assert torch.equal(ops.patchify(img.to(BF16), P), rows)

    x = _rand((B, 5, 768), cuda_device, 1.0, 40)
    gm = ops.group_mean(x, 32)
    ref = rearrange(x.float(), "b n (c h) -> b n c h", c=32).mean(-1)
    assert (gm.float() - ref).abs().max() < 1e-2

    y = ops.affine(x, 8.094, 1.468)
    assert (y.float() - (x.float() * 8.094 + 1.468)).abs().max() < 0.07

    lat = _rand((B, 7, 32), cuda_device, 1.0, 41)
    Wp = _rand((1024, 32), cuda_device, 0.2, 42)
    bp = _rand((1024,), cuda_device, 0.2, 43)
    got = ops.inproj_repeat(lat, Wp, bp)
    ref = (lat.float() @ Wp.float().t() + bp.float()).to(BF16).float() + lat.float().repeat_interleave(32, dim=-1)
    assert (got.float() - ref).abs().max() < 3e-2

    s2p = _rand((B, g * g, 4 * 1024), cuda_device, 1.0, 44)
    got = ops.pixel_shuffle(s2p, g, 2, 1024)
    ref = rearrange(s2p, "b (h w) (x y c) -> b (h x w y) c", h=g, w=g, x=2, y=2)
    assert torch.equal(got, ref)

    tok = _rand((B, g * g, 16 * 16 * 3), cuda_device, 1.0, 45)
    got = ops.unpatchify_clamp(tok, g, 16, torch.float32)
    ref = torch.einsum("nhwpqc->nchpwq", tok.float().view(B, g, g, 16, 16, 3)).reshape(B, 3, g * 16, g * 16)
    assert torch.equal(got, ref.clamp(-1, 1))

    xx = torch.zeros((B, 5, 768), dtype=BF16, device=cuda_device)
    cls, pos = _rand((768,), cuda_device, 1.0, 46), _rand((768,), cuda_device, 1.0, 47)
    ops.fill_cls_row(xx, cls, pos)
    assert torch.equal(xx[:, 4], (cls.float() + pos.float()).to(BF16).expand(B, -1))
    assert (xx[:, :4] == 0).all()


@pytest.mark.parametrize("M,N,K", [(4160, 2304, 768), (300, 1000, 264), (16384, 1024, 1024)])
@pytest.mark.parametrize("epi", ["bias", "gelu", "swiglu"])
def test_gemm_layernorm_fold(cuda_device, tile, M, N, K, epi):
    """LayerNorm folded into the GEMM epilogue == F.layer_norm followed by the Linear (+ activation)."""
    from ming_univision_b200 import ops

    if tile in ("pair128", "single128", "single128-direct") and epi == "swiglu":
        pytest.skip("SwiGLU uses 256-wide tiles")
    x = (_rand((M, K), cuda_device, 1.5, 50).float() + 0.7).to(BF16)  # non-zero mean exercises the mean * csum term
    gamma = (_rand((K,), cuda_device, 0.1, 51).float() + 1).to(BF16)
    beta = _rand((K,), cuda_device, 0.1, 52)
    ln = F.layer_norm(x.float(), (K,), gamma.float(), beta.float(), 1e-6)
    stats = ops.row_stats(x)
    assert torch.allclose(stats[:, 0, 0], x.float().sum(1), rtol=1e-5, atol=1e-2)
    assert torch.allclose(stats[:, 0, 1], (x.float() ** 2).sum(1), rtol=1e-5, atol=1e-2)
    if epi == "swiglu":
        H = 344 if N == 1000 else N // 2
        w12 = _rand((2 * H, K), cuda_device, 1.0 / math.sqrt(K), 53)
        b12 = _rand((2 * H,), cuda_device, 0.2, 54)
        wf, cs, bf = ops.fold_layernorm(w12, b12, gamma, beta)
        wp, _, Hp = ops.pack_swiglu(wf, None)
        out = ops.linear(x, wp, None, epi=ops.EPI_SWIGLU,
                         ln_fold=(stats, ops.pack_swiglu_f32(cs, H, Hp), ops.pack_swiglu_f32(bf, H, Hp), 1e-6))
        x12 = ln @ w12.float().t() + b12.float()
        ref = F.silu(x12[:, :H]) * x12[:, H:]
        got = out[:, :H]
        assert (out[:, H:].float() == 0).all()
    else:
        w = _rand((N, K), cuda_device, 1.0 / math.sqrt(K), 55)
        b = _rand((N,), cuda_device, 0.5, 56)
        wf, cs, bf = ops.fold_layernorm(w, b, gamma, beta)
        out = ops.linear(x, wf, None, epi=ops.EPI_GELU if epi == "gelu" else ops.EPI_BIAS, ln_fold=(stats, cs, bf, 1e-6))
        ref = ln @ w.float().t() + b.float()
        if epi == "gelu":
            ref = F.gelu(ref)
        got = out
    # bf16 operands (x and gamma-scaled W) + bf16 output: same error budget as LN -> bf16 -> GEMM
    assert _rel_err(got, ref) < 1e-2
    assert ((got.float() - ref).abs() <= 3e-2 * ref.abs() + 5e-2).all()


def test_gemm_residual_row_stats(cuda_device, tile):
    """The RESIDUAL epilogue leaves (sum, sum of squares) of every row it wrote — the input of the next folded GEMM."""
    from ming_univision_b200 import ops

    if tile.endswith("-direct"):
        pytest.skip("row statistics come from the staged (TMA) epilogue")
    M, N, K = 4160, 1024, 1024
    x = _rand((M, K), cuda_device, 1.0, 60)
    w = _rand((N, K), cuda_device, 1.0 / math.sqrt(K), 61)
    b = _rand((N,), cuda_device, 0.5, 62)
    stream = _rand((M, N), cuda_device, 1.0, 63)
    stats = torch.full((M, N // 64, 2), 123.0, dtype=torch.float32, device=cuda_device)
    ops.linear(x, w, b, epi=ops.EPI_RESIDUAL, residual=stream, out=stream, stats_out=stats)
    boxes = stream.float().view(M, N // 64, 64)
    assert torch.allclose(stats[:, :, 0], boxes.sum(-1), rtol=1e-4, atol=1e-2)
    assert torch.allclose(stats[:, :, 1], (boxes ** 2).sum(-1), rtol=1e-4, atol=1e-2)
    # and it is bitwise reproducible (one writer per slot, no atomics)
    stream2 = _rand((M, N), cuda_device, 1.0, 63)
    stats2 = torch.empty_like(stats)
    ops.linear(x, w, b, epi=ops.EPI_RESIDUAL, residual=stream2, out=stream2, stats_out=stats2)
    assert torch.equal(stats, stats2) and torch.equal(stream, stream2)
