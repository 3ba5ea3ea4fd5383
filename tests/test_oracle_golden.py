"""Pins oracle/mingtok_oracle.py against outputs of the UNMODIFIED reference (tests/golden/*.npz, produced by
tests/golden/make_golden.py from /root/reference).  CPU only."""
import os

import numpy as np
import pytest
import torch

from ming_univision_b200 import synthetic
from oracle import mingtok_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    return np.load(os.path.join(GOLD, name))


@pytest.mark.parametrize("size", [128, 64])
def test_tiny_mingtok_matches_reference(size):
    g = _load(f"mingtok_tiny_{size}.npz")
    cfg = synthetic.MINGTOK_TINY_CONFIG
    sd = synthetic.mingtok_state_dict(cfg, int(g["seed"]))
    img = synthetic.synthetic_images(int(g["batch"]), size, seed=int(g["img_seed"]))
    with torch.no_grad():
        out = O.mingtok_forward(sd, img, cfg)
        recon = O.pixel_decoder_forward(sd, out["x_norm_patchtokens"], cfg["semantic_decoder"], cfg["pixel_decoder"])
    for key, got in (("latent", out["latent"]), ("feats", out["x_norm_patchtokens"]), ("recon", recon)):
        ref = torch.from_numpy(g[key])
        assert got.shape == ref.shape
        assert torch.allclose(got, ref, atol=2e-5, rtol=1e-5), f"{key}: max diff {(got - ref).abs().max()}"
    # incremental decoding through the oracle's KV cache == the reference's DynamicCache path == the full pass
    caches = O.new_decoder_caches(sd)
    steps = g["feats_incremental"].shape[1]
    with torch.no_grad():
        inc = torch.cat([O.mingtok_forward_feature_decoder(sd, out["latent"][:, t:t + 1], cfg, caches)
                         for t in range(steps)], dim=1)
    ref = torch.from_numpy(g["feats_incremental"])
    assert torch.allclose(inc, ref, atol=2e-5, rtol=1e-5)
    assert torch.allclose(inc, out["x_norm_patchtokens"][:, :steps], atol=5e-5, rtol=1e-5)


def test_full_size_mingtok_matches_reference_samples():
    """Full-size (697.7 M parameter) model, 1x3x256x256 (BASELINE config 1): strided samples of the reference run."""
    g = _load("mingtok_full_256.npz")
    cfg = synthetic.MINGTOK_CONFIG
    sd = synthetic.mingtok_state_dict(cfg, int(g["seed"]))
    img = synthetic.synthetic_images(1, 256, seed=int(g["img_seed"]))
    torch.set_num_threads(os.cpu_count())
    with torch.no_grad():
        out = O.mingtok_forward(sd, img, cfg)
        recon = O.pixel_decoder_forward(sd, out["x_norm_patchtokens"], cfg["semantic_decoder"], cfg["pixel_decoder"])
        caches = O.new_decoder_caches(sd)
        inc = torch.cat([O.mingtok_forward_feature_decoder(sd, out["latent"][:, t:t + 1], cfg, caches)
                         for t in range(4)], dim=1)
    for key, got in (("latent", out["latent"]), ("feats", out["x_norm_patchtokens"]), ("recon", recon),
                     ("feats_incremental", inc)):
        flat = got.flatten()
        ref = torch.from_numpy(g[key + "_val"])
        sel = flat[torch.from_numpy(g[key + "_idx"])]
        assert torch.allclose(sel, ref, atol=2e-3, rtol=1e-3), f"{key}: max diff {(sel - ref).abs().max()}"
        mean, std, amax = g[key + "_stats"]
        assert abs(flat.mean().item() - mean) < 1e-3 + 1e-3 * abs(mean)
        assert abs(flat.std().item() - std) < 1e-3 * std + 1e-4


def test_param_count_matches_survey():
    """697.7 M parameters."""
    shapes = synthetic.mingtok_param_shapes(synthetic.MINGTOK_CONFIG)
    n = sum(int(np.prod(s)) for s in shapes.values())
    assert abs(n - 697.7e6) < 0.1e6, n


# ---------------------------------------------------------------------------------------------------------------
# rectified-flow head
# ---------------------------------------------------------------------------------------------------------------
def test_rf_tiny_matches_reference():
    from oracle import rf_oracle as R

    g = _load("rf_tiny.npz")
    cfg = synthetic.RF_TINY_CONFIG
    sd = synthetic.rf_state_dict(cfg, int(g["seed"]))
    with torch.no_grad():
        v = R.net_forward(sd, torch.from_numpy(g["net_x"]), torch.from_numpy(g["net_t"]), torch.from_numpy(g["net_c"]))
    assert torch.allclose(v, torch.from_numpy(g["net_v"]), atol=2e-5, rtol=1e-5)
    for B in (1, 2, 3):
        tc, ic, temp = (float(x) for x in g[f"B{B}_cfg"])
        with torch.no_grad():
            x = R.sample(sd, torch.from_numpy(g[f"B{B}_z"]), torch.from_numpy(g[f"B{B}_noise"]),
                         cfg["num_sampling_steps"], temp, tc, ic)
        ref = torch.from_numpy(g[f"B{B}_x"])
        assert torch.allclose(x, ref, atol=5e-5, rtol=1e-5), f"B={B}: {(x - ref).abs().max()}"


def test_rf_full_matches_reference():
    """Default-size head (width 3072, depth 12, mult 4, 16 steps; 1.285 B parameters) against the reference run."""
    from oracle import rf_oracle as R

    g = _load("rf_full.npz")
    cfg = synthetic.RF_CONFIG
    shapes = synthetic.rf_param_shapes(cfg)
    assert abs(sum(int(np.prod(s)) for s in shapes.values()) - 1.285e9) < 2e6
    sd = synthetic.rf_state_dict(cfg, int(g["seed"]))
    torch.set_num_threads(os.cpu_count())
    B = 2
    tc, ic, temp = (float(x) for x in g[f"B{B}_cfg"])
    with torch.no_grad():
        x = R.sample(sd, torch.from_numpy(g[f"B{B}_z"]), torch.from_numpy(g[f"B{B}_noise"]), 16, temp, tc, ic)
    ref = torch.from_numpy(g[f"B{B}_x"])
    assert torch.allclose(x, ref, atol=2e-3, rtol=1e-3), f"{(x - ref).abs().max()}"
