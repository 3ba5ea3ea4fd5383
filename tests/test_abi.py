"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/mingb200.h
declares, the ctypes table covers all of them, and the product path refuses to run without a GPU (no fallback)."""
import ctypes
import os

import pytest
import torch

from ming_univision_b200 import _lib, synthetic


def test_library_exports_every_header_symbol():
    lib = _lib.load()
    names = _lib.header_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/mingb200.h but not exported by libmingb200.so"
    undeclared = set(_lib.SIGNATURES) - set(names)
    assert not undeclared, f"ctypes table binds symbols missing from the header: {undeclared}"
    unbound = set(names) - set(_lib.SIGNATURES) - {"mb_last_error"}
    assert not unbound, f"header symbols without a ctypes signature: {unbound}"
    assert lib.mb_abi_version() >= 1
    assert isinstance(lib.mb_last_error(), bytes)


def test_no_cpu_fallback():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = _lib.load()
    assert lib.mb_device_ok() == 0
    rc = lib.mb_gemm_bf16(None, 8, None, 8, None, None, 8, 1, 8, 8, 0, None, 0, 0, 0, 0, None)
    assert rc == -3  # MB_ERR_ARCH
    assert b"sm_100" in lib.mb_last_error()
    with pytest.raises(RuntimeError):
        _lib.require_device()
    from ming_univision_b200.mingtok import MingTok, MingTokConfig

    m = MingTok(MingTokConfig(**synthetic.MINGTOK_TINY_CONFIG))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.forward(torch.zeros(1, 3, 128, 128))


def test_state_dict_schema_matches_reference_keys():
    """Same keys and shapes as the reference's MingTok.state_dict() and HF save/load round trip."""
    import tempfile

    from ming_univision_b200.mingtok import MingTok, MingTokConfig

    cfg = synthetic.MINGTOK_TINY_CONFIG
    m = MingTok(MingTokConfig(**cfg))
    shapes = synthetic.mingtok_param_shapes(cfg)
    sd = m.state_dict()
    assert set(sd) == set(shapes)
    for k, v in sd.items():
        assert tuple(v.shape) == tuple(shapes[k]), k
    m.load_state_dict(synthetic.mingtok_state_dict(cfg, 3), strict=True)
    with tempfile.TemporaryDirectory() as d:
        m.save_pretrained(d)
        m2 = MingTok.from_pretrained(d)
    assert all(torch.equal(a, b) for a, b in zip(m.state_dict().values(), m2.state_dict().values()))
    assert (m.latent_dim, m.feature_dim, m.patch_size) == (32, 128, 32)


def test_reference_keys_live():
    """When the reference tree is present (build container), compare the key schema against the real thing."""
    from oracle import ref_shims

    if not ref_shims.reference_available():
        pytest.skip("reference tree not present")
    cfg = synthetic.MINGTOK_TINY_CONFIG
    ref = ref_shims.build_reference_mingtok(cfg)
    shapes = synthetic.mingtok_param_shapes(cfg)
    ref_sd = ref.state_dict()
    assert set(ref_sd) == set(shapes)
    assert all(tuple(ref_sd[k].shape) == tuple(shapes[k]) for k in shapes)


def test_rf_state_dict_schema_live():
    """RectifiedFlowLoss keys/shapes == the reference module's (and == the synthetic factory's)."""
    from ming_univision_b200.diff_loss_rf_swiglu import RectifiedFlowLoss
    from oracle import ref_shims

    cfg = synthetic.RF_TINY_CONFIG
    m = RectifiedFlowLoss(cfg["target_channels"], cfg["z_channels"], cfg["depth"], cfg["width"],
                          str(cfg["num_sampling_steps"]), mlp_mult=cfg["mlp_mult"])
    shapes = synthetic.rf_param_shapes(cfg)
    sd = m.state_dict()
    assert set(sd) == set(shapes) and all(tuple(sd[k].shape) == tuple(shapes[k]) for k in shapes)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.sample(torch.zeros(2, cfg["z_channels"]), text_cfg=3.0)
    if not ref_shims.reference_available():
        return
    ref_shims.install()
    import contextlib
    import io

    os.environ["XFORMERS_DISABLED"] = "1"
    from diff_loss_rf_swiglu import RectifiedFlowLoss as RefRF

    with contextlib.redirect_stdout(io.StringIO()):
        ref = RefRF(cfg["target_channels"], cfg["z_channels"], cfg["depth"], cfg["width"],
                    str(cfg["num_sampling_steps"]), mlp_mult=cfg["mlp_mult"])
    rsd = ref.state_dict()
    assert set(rsd) == set(shapes) and all(tuple(rsd[k].shape) == tuple(shapes[k]) for k in shapes)
