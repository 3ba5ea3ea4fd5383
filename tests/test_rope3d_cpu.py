"""CPU checks of the 3-D multimodal RoPE variant (`rope_scaling.type == "3D"`):

1. the oracle (oracle/bailing_oracle.py: mrope_tables / apply_mrope) against the golden vectors of the unmodified
   reference functions (tests/golden/rope3d.npz) and — when the checkout is present — against those functions live;
2. the device code (ming_univision_b200/csrc/rope3d_core.h, the per-thread body of rope3d_kv_append_kernel) compiled with
   g++ and walked over the kernel's grid: cache layout, slot arithmetic, section -> component selection, the V copy, and
   the rounding chain (fp32 products and sum, ONE bf16 rounding).  cosf / sinf / powf come from the host libm here and
   from CUDA's libdevice on the GPU, so the comparison allows one bf16 ulp on a small fraction of the elements.
"""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest
import torch

from oracle import bailing_oracle as O
from oracle import ref_shims

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "rope3d.npz"))


def test_oracle_matches_reference_golden(golden):
    g = golden
    B, S, H, Hkv, hd = (int(v) for v in g["dims"])
    q, k = torch.from_numpy(g["q"]).to(torch.bfloat16), torch.from_numpy(g["k"]).to(torch.bfloat16)
    cos, sin = O.mrope_tables(hd, float(g["theta"]), torch.from_numpy(g["pos"]))
    qe, ke = O.apply_mrope(q, k, cos, sin)
    assert qe.dtype == torch.float32
    assert torch.equal(qe, torch.from_numpy(g["q_rot"])) and torch.equal(ke, torch.from_numpy(g["k_rot"]))


@pytest.mark.skipif(not ref_shims.reference_available(), reason="reference checkout not present")
def test_oracle_matches_reference_live():
    import importlib

    ref_shims.install()
    m = importlib.import_module("modeling_bailing_moe")
    gen = torch.Generator().manual_seed(5)
    B, S, H, Hkv, hd = 3, 9, 16, 4, 128
    q = torch.randn(B, H, S, hd, generator=gen).to(torch.bfloat16)
    k = torch.randn(B, Hkv, S, hd, generator=gen).to(torch.bfloat16)
    pos = torch.randint(0, 5000, (3, B, S), generator=gen)
    rot = m.BailingMoe3DRotaryEmbedding(hd, max_position_embeddings=4096, base=600000.0)
    cos, sin = rot(k, position_ids=pos)
    qr, kr = m.apply_multimodal_rotary_pos_emb(q, k, cos, sin)
    c2, s2 = O.mrope_tables(hd, 600000.0, pos)
    qo, ko = O.apply_mrope(q, k, c2, s2)
    assert torch.equal(cos, c2) and torch.equal(sin, s2) and torch.equal(qr, qo) and torch.equal(kr, ko)


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    so = tmp_path_factory.mktemp("rope3d") / "librope3d_emu.so"
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-shared", "-fPIC", "-Wall", "-Werror", "-o",
                    str(so), os.path.join(ROOT, "tests", "native", "rope3d_emu.cpp")], check=True)
    lib = C.CDLL(str(so))
    lib.emu_float_to_bf16_bits.restype = C.c_uint
    lib.emu_float_to_bf16_bits.argtypes = [C.c_float]
    return lib


def bits(t: torch.Tensor) -> np.ndarray:
    return np.ascontiguousarray(t.contiguous().view(torch.int16).numpy().view(np.uint16))


def test_bf16_rounding_helper(emu):
    """float_to_bf16_bits == torch's round-to-nearest-even cast, including ties, subnormals, infinities."""
    vals = torch.tensor([0.0, -0.0, 1.0, 1.00390625, 1.005859375, 1.001953125, -3.140625, 65504.0, 3.3895e38, 1e-40,
                         float("inf"), -float("inf")], dtype=torch.float32)
    vals = torch.cat([vals, torch.randn(2000, generator=torch.Generator().manual_seed(1)) * 37.0])
    want = bits(vals.to(torch.bfloat16))
    got = np.array([emu.emu_float_to_bf16_bits(float(v)) for v in vals.tolist()], dtype=np.uint16)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("B,S,H,Hkv,t0,Tmax", [(2, 5, 4, 2, 0, 16), (1, 1, 16, 4, 37, 64), (3, 1, 16, 4, 11, 32),
                                               (1, 12, 16, 4, 3, 20)])
def test_emulated_kernel_vs_oracle(emu, B, S, H, Hkv, t0, Tmax):
    hd, theta, sec = 128, 600000.0, (16, 24, 24)
    gen = torch.Generator().manual_seed(B * 100 + S)
    qkv = torch.randn(B * S, (H + 2 * Hkv) * hd, generator=gen).to(torch.bfloat16)
    pos = torch.stack([torch.randint(0, 3000, (B, S), generator=gen), torch.randint(0, 70, (B, S), generator=gen),
                       torch.randint(0, 70, (B, S), generator=gen)])
    # reference dataflow (:876-893): split heads, [B, heads, S, hd], rotate, bf16 once
    x = qkv.view(B, S, H + 2 * Hkv, hd)
    q, k, v = x[:, :, :H].transpose(1, 2), x[:, :, H:H + Hkv].transpose(1, 2), x[:, :, H + Hkv:].transpose(1, 2)
    cos, sin = O.mrope_tables(hd, theta, pos)
    qe, ke = O.apply_mrope(q, k, cos, sin, sec)
    want_q = qe.to(torch.bfloat16).transpose(1, 2).reshape(B * S, H * hd)
    want_k, want_v = ke.to(torch.bfloat16), v

    kc = torch.full((B, Hkv, Tmax, hd), 7.0).to(torch.bfloat16)
    vc = torch.full((B, Hkv, Tmax, hd), 7.0).to(torch.bfloat16)
    q_out = np.zeros((B * S, H * hd), dtype=np.uint16)
    kcb, vcb = bits(kc), bits(vc)
    pos_i32 = np.ascontiguousarray(pos.reshape(3, -1).numpy().astype(np.int32))
    emu.emu_rope3d_kv_append(bits(qkv).ctypes.data_as(C.c_void_p), pos_i32.ctypes.data_as(C.c_void_p),
                             q_out.ctypes.data_as(C.c_void_p), kcb.ctypes.data_as(C.c_void_p),
                             vcb.ctypes.data_as(C.c_void_p), B, S, H, Hkv, hd, Tmax, t0, C.c_float(theta), sec[0], sec[1])
    to_f = lambda a: torch.from_numpy(a.view(np.int16).copy()).view(torch.bfloat16).float()  # noqa: E731
    got_q, got_k, got_v = to_f(q_out), to_f(kcb), to_f(vcb)

    # Tolerance.  The angle pos * inv_freq is an fp32 number of size up to ~1e3 rad, so ONE ulp of difference in inv_freq
    # (host libm powf vs torch's vectorised pow vs CUDA's powf) moves it by up to ang * 2^-23 rad, and the result by that
    # times (|x1| + |x2|) — an ABSOLUTE error that exceeds a bf16 ulp where the two products cancel.  The reference's
    # own CPU and GPU runs differ from each other in exactly this way.  Allowed: one bf16 ulp + 4 ulps of the angle.
    inv_freq = 1.0 / (theta ** (torch.arange(0, hd, 2).float() / hd))
    comp = torch.tensor([0 if i < sec[0] else (1 if i < sec[0] + sec[1] else 2) for i in range(hd // 2)])
    ang = (pos[..., None].float() * inv_freq)[comp, :, :, torch.arange(hd // 2)].permute(1, 2, 0)  # [B, S, hd/2]
    ang = torch.cat([ang, ang], dim=-1)[:, None]                                                   # [B, 1, S, hd]

    def close(got, want, x, what):
        diff = (got - want.float()).abs()
        big = torch.maximum(torch.maximum(got.abs(), want.float().abs()), torch.tensor(2.0 ** -100))
        ulp = torch.exp2(torch.floor(torch.log2(big)) - 7)  # bf16 spacing in the binade of the larger value
        mag = x.float().abs() + O.rotate_half(x.float()).abs()
        tol = ulp + mag * (ang + 1.0) * 2.0 ** -21
        assert bool((diff <= tol).all()), (what, float((diff / tol).max()))
        assert float((diff > 0).float().mean()) < 0.05, (what, float((diff > 0).float().mean()))

    close(got_q.view(B, S, H, hd).transpose(1, 2), want_q.view(B, S, H, hd).transpose(1, 2), q, "q")
    close(got_k[:, :, t0:t0 + S], want_k, k, "k")
    assert torch.equal(got_v[:, :, t0:t0 + S], want_v.float()), "v is a plain copy"
    # nothing outside the appended slots was touched
    untouched = torch.ones(Tmax, dtype=torch.bool)
    untouched[t0:t0 + S] = False
    assert bool((got_k[:, :, untouched] == 7.0).all()) and bool((got_v[:, :, untouched] == 7.0).all())


def test_emulated_kernel_vs_reference_golden(emu, golden):
    g = golden
    B, S, H, Hkv, hd = (int(v) for v in g["dims"])
    q, k = torch.from_numpy(g["q"]).to(torch.bfloat16), torch.from_numpy(g["k"]).to(torch.bfloat16)
    v = torch.zeros_like(k)
    qkv = torch.cat([q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2)], dim=2).reshape(B * S, -1)
    kcb, vcb = np.zeros((B, Hkv, S, hd), dtype=np.uint16), np.zeros((B, Hkv, S, hd), dtype=np.uint16)
    q_out = np.zeros((B * S, H * hd), dtype=np.uint16)
    pos_i32 = np.ascontiguousarray(g["pos"].reshape(3, -1).astype(np.int32))
    emu.emu_rope3d_kv_append(bits(qkv).ctypes.data_as(C.c_void_p), pos_i32.ctypes.data_as(C.c_void_p),
                             q_out.ctypes.data_as(C.c_void_p), kcb.ctypes.data_as(C.c_void_p),
                             vcb.ctypes.data_as(C.c_void_p), B, S, H, Hkv, hd, S, 0, C.c_float(float(g["theta"])), 16, 24)
    got_q = torch.from_numpy(q_out.view(np.int16).copy()).view(torch.bfloat16).float().view(B, S, H, hd).transpose(1, 2)
    got_k = torch.from_numpy(kcb.view(np.int16).copy()).view(torch.bfloat16).float()
    for got, want in ((got_q, torch.from_numpy(g["q_rot"])), (got_k, torch.from_numpy(g["k_rot"]))):
        want16 = want.to(torch.bfloat16).float()
        assert float((got - want16).abs().max()) <= float(want16.abs().max()) * 2.0 ** -7
        assert float((got != want16).float().mean()) < 0.02


def test_section_check_through_the_c_abi():
    from ming_univision_b200 import _lib

    lib = _lib.load()
    if torch.cuda.is_available():
        pytest.skip("GPU present: argument checks run behind the device check")
    assert lib.mb_rope3d_kv_append(None, None, None, None, None, 1, 1, 16, 4, 128, 8, None, 0, 6e5, 16, 24, 24, None) == -3
