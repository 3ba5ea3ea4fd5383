"""Rectified-flow head on the GPU: the weight-streaming kernels against PyTorch fp32 references of the same ops, and
`RectifiedFlowLoss.sample` (product path) against the fp32 oracle / the golden outputs of the unmodified reference.

Stated tolerance for sample(): relative L2 <= 4e-2 on the final latent after 16 Euler steps x 12 bf16 res-blocks
(the reference itself runs this head in bf16 under autocast; the oracle and the golden vectors are fp32)."""
import math
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from ming_univision_b200 import synthetic
from parity_metrics import rel_l2

pytestmark = pytest.mark.gpu
BF16 = torch.bfloat16
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _rand(shape, dev, scale=1.0, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(dev).to(BF16)


@pytest.mark.parametrize("M,N,K", [(1, 3072, 3072), (2, 3072, 8192), (3, 16384, 3072), (3, 32, 3072), (3, 3072, 32),
                                   (8, 1000, 264), (5, 2048, 1024), (3, 126464, 2048)])
@pytest.mark.parametrize("epi", ["bias", "gelu", "silu", "residual", "gated"])
def test_gemv_epilogues(cuda_device, M, N, K, epi):
    from ming_univision_b200 import ops

    if N > 100000 and epi != "bias":
        pytest.skip("vocab-sized N only with the plain epilogue")
    x = _rand((M, K), cuda_device, 1.0, 1)
    w = _rand((N, K), cuda_device, 1.0 / math.sqrt(K), 2)
    b = _rand((N,), cuda_device, 0.5, 3)
    r = _rand((M, N), cuda_device, 1.0, 4)
    gt = _rand((M, N), cuda_device, 1.0, 5)
    pre = x.float() @ w.float().t() + b.float()
    f32 = None
    if epi == "bias":
        f32 = torch.empty((M, N), dtype=torch.float32, device=cuda_device)
        out = ops.gemv(x, w, b, out_f32=f32)
        ref = pre
    elif epi == "gelu":
        out, ref = ops.gemv(x, w, b, epi=ops.EPI_GELU), F.gelu(pre.to(BF16).float())
    elif epi == "silu":
        out, ref = ops.gemv(x, w, b, epi=ops.EPI_SILU), F.silu(pre.to(BF16).float())
    elif epi == "residual":
        out, ref = ops.gemv(x, w, b, epi=ops.EPI_RESIDUAL, residual=r), pre.to(BF16).float() + r.float()
    else:
        out = ops.gemv(x, w, b, epi=ops.EPI_GATED, residual=r, gate=gt)
        ref = r.float() + (gt.float() * pre.to(BF16).float()).to(BF16).float()
    err = (out.float() - ref).abs()
    tol = 2.0 ** -7 * ref.abs() + 2.0 ** -7 * pre.abs() * (2.0 if epi == "gated" else 1.0) + 2e-3
    assert (err <= tol).all(), f"max err {err.max().item()}"
    if f32 is not None:
        assert torch.equal(f32, out.float())


@pytest.mark.parametrize("M,K,H", [(3, 3072, 8192), (2, 1024, 2736), (1, 128, 344)])
def test_gemv_swiglu(cuda_device, M, K, H):
    from ming_univision_b200 import ops

    x = _rand((M, K), cuda_device, 1.0, 7)
    w12 = _rand((2 * H, K), cuda_device, 1.0 / math.sqrt(K), 8)
    b12 = _rand((2 * H,), cuda_device, 0.2, 9)
    out = ops.gemv(x, w12, b12, epi=ops.EPI_SWIGLU)
    x12 = (x.float() @ w12.float().t() + b12.float()).to(BF16).float()
    ref = F.silu(x12[:, :H]).to(BF16).float() * x12[:, H:]
    assert out.shape == (M, H)
    assert ((out.float() - ref).abs() <= 2.0 ** -6 * ref.abs() + 5e-3).all()


def test_rf_row_helpers(cuda_device):
    from ming_univision_b200 import ops

    B, D = 3, 3072
    x = _rand((B, D), cuda_device, 2.0, 10)
    g = (_rand((D,), cuda_device, 0.1, 11).float() + 1).to(BF16)
    b = _rand((D,), cuda_device, 0.1, 12)
    mod = _rand((B, 3 * D), cuda_device, 0.5, 13)
    y = ops.adaln_modulate(x, g, b, mod[:, :D], mod[:, D:2 * D])
    ref = F.layer_norm(x.float(), (D,), g.float(), b.float(), 1e-6) * (1 + mod[:, D:2 * D].float()).to(BF16).float() \
        + mod[:, :D].float()
    assert ((y.float() - ref).abs() <= 2.0 ** -7 * ref.abs() + 2e-3).all()
    y2 = ops.adaln_modulate(x, None, None, mod[:, :D], mod[:, D:2 * D])
    ref2 = F.layer_norm(x.float(), (D,), None, None, 1e-6) * (1 + mod[:, D:2 * D].float()).to(BF16).float() \
        + mod[:, :D].float()
    assert ((y2.float() - ref2).abs() <= 2.0 ** -7 * ref2.abs() + 2e-3).all()

    temb, c = _rand((16, D), cuda_device, 1.0, 14), _rand((B, D), cuda_device, 1.0, 15)
    sy = ops.silu_add_rows(temb, c)
    ref = F.silu((temb.float()[:, None] + c.float()[None]).to(BF16).float()).reshape(16 * B, D)
    assert ((sy.float() - ref).abs() <= 2.0 ** -7 * ref.abs() + 1e-3).all()

    for Bc in (1, 2, 3):
        v = _rand((Bc, 32), cuda_device, 1.0, 16 + Bc)
        x0 = torch.randn((1, 32), device=cuda_device).repeat(Bc, 1).contiguous()
        xf, xb = x0.clone(), torch.zeros((Bc, 32), dtype=BF16, device=cuda_device)
        ops.rf_euler_step(xf, xb, v, 1 / 16, 3.0, 1.1)  # one sample of Bc CFG rows
        vf = v.float()
        if Bc == 3:
            vg = vf[1] + 1.1 * (vf[2] - vf[1]) + 3.0 * (vf[0] - vf[2])
        elif Bc == 2:
            vg = vf[1] + 3.0 * (vf[0] - vf[1])
        else:
            vg = vf[0]
        ref = x0[0] + vg / 16
        assert (xf - ref[None]).abs().max() < 3e-3 * (1 + vg.abs().max())
        assert torch.equal(xb.float(), xf.to(BF16).float())
        assert all(torch.equal(xf[0], xf[i]) for i in range(Bc))


def _build_rf(cfg, sd, device):
    from ming_univision_b200.diff_loss_rf_swiglu import RectifiedFlowLoss

    with torch.device(device):
        m = RectifiedFlowLoss(target_channels=cfg["target_channels"], z_channels=cfg["z_channels"], depth=cfg["depth"],
                              width=cfg["width"], num_sampling_steps=str(cfg["num_sampling_steps"]),
                              mlp_mult=cfg["mlp_mult"])
    m.load_state_dict({k: v.to(device) for k, v in sd.items()}, strict=True)
    return m.to(BF16)


@pytest.mark.parametrize("name", ["tiny", "full"])
def test_rf_sample_vs_reference_golden(cuda_device, name):
    cfg = synthetic.RF_TINY_CONFIG if name == "tiny" else synthetic.RF_CONFIG
    g = np.load(os.path.join(GOLD, f"rf_{name}.npz"))
    m = _build_rf(cfg, synthetic.rf_state_dict(cfg, int(g["seed"])), cuda_device)
    for B in (1, 2, 3):
        if f"B{B}_x" not in g:
            continue
        tc, ic, temp = (float(x) for x in g[f"B{B}_cfg"])
        z = torch.from_numpy(g[f"B{B}_z"]).to(cuda_device)
        noise = torch.from_numpy(g[f"B{B}_noise"]).to(cuda_device)
        ref = torch.from_numpy(g[f"B{B}_x"])
        x = m.sample(z, temperature=temp, text_cfg=tc, image_cfg=ic, noise=noise)
        assert x.dtype == torch.float32 and x.shape == ref.shape
        assert all(torch.equal(x[0], x[i]) for i in range(B)) or B == 1 or tc == 1.0
        e = rel_l2(x, ref)
        print(f"rf {name} B={B}: rel-L2 vs reference(fp32) {e:.3e}")
        assert e < 4e-2
        # CUDA-graph replay == eager launches, bit for bit
        m.use_cuda_graph = False
        x2 = m.sample(z, temperature=temp, text_cfg=tc, image_cfg=ic, noise=noise)
        m.use_cuda_graph = True
        assert torch.equal(x, x2)
        # second replay of the cached graph with different inputs
        x3 = m.sample(z * 0.5, temperature=temp, text_cfg=tc, image_cfg=ic, noise=noise)
        assert not torch.equal(x3, x)
        assert torch.equal(m.sample(z, temperature=temp, text_cfg=tc, image_cfg=ic, noise=noise), x)


def test_rf_persistent_sampler_vs_layer_path(cuda_device, monkeypatch):
    """The persistent sampler kernel (csrc/rf_fused.cu: whole Euler loop in one launch, per-CTA packed weights streamed by
    cp.async.bulk) against the launch-per-layer path (MB_RF_FUSED=0) on the default-size head: the same rounding points
    with another summation order, so the samples agree far inside their common distance from the fp32 reference
    (1.3e-2); both must leave identical CFG rows, and the fused kernel must be bit-reproducible."""
    from ming_univision_b200 import ops  # noqa: F401

    cfg = synthetic.RF_CONFIG
    g = np.load(os.path.join(GOLD, "rf_full.npz"))
    m = _build_rf(cfg, synthetic.rf_state_dict(cfg, int(g["seed"])), cuda_device)
    for B in (1, 2, 3):
        z = (torch.randn((B, cfg["z_channels"]), generator=torch.Generator().manual_seed(5 + B))).to(cuda_device)
        noise = torch.randn((1, 32), generator=torch.Generator().manual_seed(9)).to(cuda_device)
        kw = dict(temperature=0.9, text_cfg=3.0, image_cfg=1.1, noise=noise)
        monkeypatch.setenv("MB_RF_FUSED", "1")
        m._graphs = {}
        a = m.sample(z, **kw)
        a2 = m.sample(z, **kw)
        m.use_cuda_graph = False
        a3 = m.sample(z, **kw)
        m.use_cuda_graph = True
        monkeypatch.setenv("MB_RF_FUSED", "0")
        m._graphs = {}
        b = m.sample(z, **kw)
        m._graphs = {}
        e = rel_l2(a, b)
        print(f"rf persistent vs layer path B={B}: rel-L2 {e:.3e}")
        assert torch.equal(a, a2) and torch.equal(a, a3)
        assert all(torch.equal(a[0], a[i]) for i in range(B))
        assert e < 1e-2, e
    monkeypatch.setenv("MB_RF_FUSED", "1")


@pytest.mark.parametrize("M", [1, 3, 8])
@pytest.mark.parametrize("norm", ["adaln", "adaln_noaffine", "rms"])
def test_gemv_fused_norm_equals_two_kernels(cuda_device, M, norm):
    """mb_gemv_bf16_norm (normalisation fused into the staging of the activation rows) against the separate row kernel
    followed by the plain streaming GEMM: same rounding points, so the outputs agree to one bf16 ulp of the accumulated
    sums (the row statistics are reduced in a different order)."""
    from ming_univision_b200 import ops

    K, N = 3072, 640
    g = torch.Generator().manual_seed(70 + M)
    rnd = lambda *sh, s=1.0: (torch.randn(sh, generator=g) * s).to(cuda_device).to(torch.bfloat16)  # noqa: E731
    x, w, b = rnd(M, K, s=2.0), rnd(N, K, s=K ** -0.5), rnd(N)
    gamma, beta = (rnd(K, s=0.1).float() + 1).to(torch.bfloat16), rnd(K, s=0.1)
    mod = rnd(M, 2 * K + 8, s=0.3)
    if norm == "rms":
        a = ops.rmsnorm(x, gamma, 1e-5)
        fused = ops.gemv_norm(x, w, b, norm="rms", gamma=gamma, eps=1e-5)
    elif norm == "adaln":
        a = ops.adaln_modulate(x, gamma, beta, mod[:, :K], mod[:, K:2 * K])
        fused = ops.gemv_norm(x, w, b, norm="adaln", gamma=gamma, beta=beta, shift=mod[:, :K], scale=mod[:, K:2 * K])
    else:
        a = ops.adaln_modulate(x, None, None, mod[:, :K], mod[:, K:2 * K])
        fused = ops.gemv_norm(x, w, b, norm="adaln", shift=mod[:, :K], scale=mod[:, K:2 * K])
    two = ops.gemv(a, w, b)
    assert fused.shape == two.shape
    err = (fused.float() - two.float()).abs()
    assert (err <= 2.0 ** -6 * two.float().abs() + 2e-2).all(), float(err.max())
    assert float((fused.float() - two.float()).norm() / two.float().norm()) < 3e-3
    # fused SwiGLU variant (the RF ResBlock form)
    w12, b12 = rnd(2 * 256, K, s=K ** -0.5), rnd(2 * 256)
    if norm != "rms":
        gm, bt = (gamma, beta) if norm == "adaln" else (None, None)
        f2 = ops.gemv_norm(x, w12, b12, norm="adaln", gamma=gm, beta=bt, shift=mod[:, :K], scale=mod[:, K:2 * K],
                           epi=ops.EPI_SWIGLU)
        t2 = ops.gemv(a, w12, b12, epi=ops.EPI_SWIGLU)
        assert float((f2.float() - t2.float()).norm() / t2.float().norm()) < 5e-3
