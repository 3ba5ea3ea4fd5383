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


def test_upsampled_position_table_matches_reference():
    """Inputs LARGER than the native resolution (the understanding path: 1024 x 1024 on a 16 x 16 table,
    vision_transformer.py:183-215): the tiny model at 256 px (4 x 4 -> 8 x 8), full tensors, and the full-size encoder +
    semantic decoder at 1 x 3 x 1024 x 1024 (S = 1025), strided samples (tests/golden/make_golden_upsample.py)."""
    g = _load("mingtok_tiny_256.npz")
    cfg = synthetic.MINGTOK_TINY_CONFIG
    sd = synthetic.mingtok_state_dict(cfg, int(g["seed"]))
    img = synthetic.synthetic_images(1, 256, seed=int(g["img_seed"]))
    with torch.no_grad():
        out = O.mingtok_forward(sd, img, cfg)
        recon = O.pixel_decoder_forward(sd, out["x_norm_patchtokens"], cfg["semantic_decoder"], cfg["pixel_decoder"])
    for key, got in (("latent", out["latent"]), ("feats", out["x_norm_patchtokens"]), ("recon", recon)):
        ref = torch.from_numpy(g[key])
        assert got.shape == ref.shape
        assert torch.allclose(got, ref, atol=2e-5, rtol=1e-5), f"{key}: max diff {(got - ref).abs().max()}"
    g = _load("mingtok_full_1024.npz")
    cfg = synthetic.MINGTOK_CONFIG
    sd = synthetic.mingtok_state_dict(cfg, int(g["seed"]))
    img = synthetic.synthetic_images(1, 1024, seed=int(g["img_seed"]))
    with torch.no_grad():
        out = O.mingtok_forward(sd, img, cfg)
    for key, got in (("latent", out["latent"]), ("feats", out["x_norm_patchtokens"])):
        assert tuple(got.shape) == tuple(g[key + "_shape"])
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
    B = 2
    tc, ic, temp = (float(x) for x in g[f"B{B}_cfg"])
    with torch.no_grad():
        x = R.sample(sd, torch.from_numpy(g[f"B{B}_z"]), torch.from_numpy(g[f"B{B}_noise"]), 16, temp, tc, ic)
    ref = torch.from_numpy(g[f"B{B}_x"])
    assert torch.allclose(x, ref, atol=2e-3, rtol=1e-3), f"{(x - ref).abs().max()}"


# ---------------------------------------------------------------------------------------------------------------
# Bailing-MoE AR path
# ---------------------------------------------------------------------------------------------------------------
def _llm_fixture():
    from oracle import bailing_oracle as L

    g = _load("llm_tiny.npz")
    cfg, vh = synthetic.LLM_TINY_CONFIG, synthetic.VISHEAD_TINY_CONFIG
    tok_cfg = synthetic.MINGTOK_TINY_CONFIG
    sd = synthetic.llm_state_dict(cfg, vh, feature_dim=tok_cfg["semantic_decoder"]["embed_dim"], seed=int(g["seed"]))
    rf_sd = {k[len("diffloss."):]: v for k, v in sd.items() if k.startswith("diffloss.")}
    return L, g, cfg, vh, tok_cfg, sd, rf_sd


def test_llm_prefill_and_cfg_step_match_reference():
    L, g, cfg, vh, tok_cfg, sd, rf_sd = _llm_fixture()
    ids = torch.from_numpy(g["prefill_ids"])
    emb = sd["model.word_embeddings.weight"][ids]
    caches = L.new_caches(cfg)
    with torch.no_grad():
        h = L.model_forward(sd, cfg, emb, torch.ones(1, ids.shape[1], dtype=torch.long), None, caches,
                            image_mask=torch.from_numpy(g["prefill_image_mask"]))
        logits = L.lm_logits(sd, h[:, -1])
    assert torch.allclose(h, torch.from_numpy(g["prefill_hidden"]), atol=2e-5, rtol=1e-5)
    assert torch.allclose(logits, torch.from_numpy(g["prefill_logits_last"]), atol=5e-5, rtol=1e-5)
    assert torch.allclose(caches[0]["k"], torch.from_numpy(g["prefill_k0"]), atol=1e-5)
    assert torch.allclose(caches[1]["v"], torch.from_numpy(g["prefill_v1"]), atol=1e-5)
    # cached decode step, 2 CFG rows, 2-D padding mask, per-row positions
    for c in caches:
        c["k"], c["v"] = c["k"].repeat(2, 1, 1, 1), c["v"].repeat(2, 1, 1, 1)
    with torch.no_grad():
        h2 = L.model_forward(sd, cfg, torch.from_numpy(g["step_x"]), torch.from_numpy(g["step_mask"]),
                             torch.from_numpy(g["step_pos"]), caches)
        z = L.vis_head(sd, h2[:, -1:])
    assert torch.allclose(h2, torch.from_numpy(g["step_hidden"]), atol=2e-5, rtol=1e-5)
    assert torch.allclose(z, torch.from_numpy(g["step_z"]), atol=5e-5, rtol=1e-5)


@pytest.mark.parametrize("name", ["t2i", "edit"])
def test_generate_image_matches_reference(name):
    """oracle generate_image (LLM step + RF sampler + MingTok cached decode + linear_proj + pixel decoder) against the
    reference's own generate_image run end to end (B = 2 and B = 3 CFG rows)."""
    L, g, cfg, vh, tok_cfg, sd, rf_sd = _llm_fixture()
    tok_sd = synthetic.mingtok_state_dict(tok_cfg, 0)
    ids = torch.from_numpy(g["prefill_ids"])
    emb = sd["model.word_embeddings.weight"][ids]
    caches = L.new_caches(cfg)
    S = ids.shape[1]
    with torch.no_grad():
        L.model_forward(sd, cfg, emb, torch.ones(1, S, dtype=torch.long), None, caches)
    start = sd["model.word_embeddings.weight"][torch.tensor([[cfg["image_start_token"]]])]

    def latent_to_sem(latent, state):
        state = O.new_decoder_caches(tok_sd) if state is None else state
        return O.mingtok_forward_feature_decoder(tok_sd, latent, tok_cfg, state), state

    noises = [torch.from_numpy(n) for n in g[f"{name}_noises"]]
    tm = torch.from_numpy(g[f"{name}_text_uncond"])
    with torch.no_grad():
        feats, lats, fmask = L.generate_image(
            sd, cfg, rf_sd, int(vh["num_sampling_steps"]), start, caches, torch.ones(1, S + 1, dtype=torch.long),
            torch.from_numpy(g[f"{name}_uncond"]), tm, latent_to_sem, lambda f: L.linear_proj(sd, f), noises,
            temperature=0.9)
        img = O.pixel_decoder_forward(tok_sd, torch.cat(feats, dim=1), tok_cfg["semantic_decoder"],
                                      tok_cfg["pixel_decoder"])
    assert torch.equal(fmask, torch.from_numpy(g[f"{name}_final_mask"]))
    assert torch.allclose(torch.cat(lats, dim=1), torch.from_numpy(g[f"{name}_latents"]), atol=2e-4, rtol=1e-4)
    assert torch.allclose(torch.cat(feats, dim=1), torch.from_numpy(g[f"{name}_feats"]), atol=5e-4, rtol=1e-4)
    assert torch.allclose(img, torch.from_numpy(g[f"{name}_image"]), atol=1e-3, rtol=1e-3)
    assert caches[0]["k"].shape[0] == int(g[f"{name}_cache_batch"]) == 1
    assert caches[0]["k"].shape[2] == int(g[f"{name}_cache_len"])


def test_moe_block_at_true_widths_matches_reference():
    """The oracle's MoE block (64 experts top-6, I = 1408, shared 2816, image gate) against the live reference's
    `BailingMoeSparseMoeBlock.forward` at the true 16B-A3B widths (tests/golden/llm_wide.npz; layer 0 only: 0.55 B
    parameters, drawn key by key in parallel)."""
    from concurrent.futures import ThreadPoolExecutor

    from oracle import bailing_oracle as L

    g = _load("llm_wide.npz")
    cfg = synthetic.LLM_WIDE_CONFIG
    pre = "model.layers.0.mlp."
    shapes = {k: v for k, v in synthetic.llm_param_shapes(dict(cfg, num_hidden_layers=1)).items() if k.startswith(pre)}
    with ThreadPoolExecutor(8) as ex:
        sd = dict(zip(shapes, ex.map(lambda kv: synthetic.llm_tensor(kv[0], kv[1], 0), shapes.items())))
    S, D = 192, cfg["hidden_size"]
    x = torch.randn((1, S, D), generator=torch.Generator().manual_seed(int(g["moe_seed"])))
    with torch.no_grad():
        y, idx = L.moe_block(sd, pre[:-1], cfg, x, torch.from_numpy(g["prefill_image_mask"]))
    rows = torch.from_numpy(g["prefill_rows"]).long()
    assert torch.equal(idx.reshape(S, -1).sort(-1).values,
                       torch.from_numpy(g["moe_topk_idx"].astype(np.int64)).sort(-1).values)
    assert torch.allclose(y[0, rows], torch.from_numpy(g["moe_y_rows"]), atol=2e-5, rtol=1e-5)
