"""Bailing-MoE AR path at the TRUE 16B-A3B widths (depth-reduced to 2 layers) on the GPU against golden
outputs of the UNMODIFIED reference (tests/golden/llm_wide.npz, tests/golden/make_golden_llm_wide.py): hidden 2048,
16 / 4 heads x 128, 64 routed experts top-6 (I = 1408) + shared expert 2816, multi_gate, vocabulary 126464, default-size
vis_head + RF head (1.285 B parameters, 16 Euler steps).

Stated tolerances (bf16 path vs the fp32 reference): hidden states / z / MoE-block output of one forward: relative L2
<= 2e-2; last-row logits: relative L2 <= 2e-2 and max-abs-diff <= 0.1 (unit-scale logits, std 1.0, max 4.5); one RF
sample from a reference-exact context (teacher forcing): latent relative L2 <= 8e-2; free-running generation: inside the
drift envelope of the reference's OWN bf16 regime (see test_wide_generate_image)."""
import os

import numpy as np
import pytest
import torch

from ming_univision_b200 import synthetic
from parity_metrics import rel_l2

pytestmark = pytest.mark.gpu
BF16 = torch.bfloat16
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "llm_wide.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


@pytest.fixture(scope="module")
def wide_model(cuda_device):
    from ming_univision_b200.mingtok import MingTokConfig
    from ming_univision_b200.modeling_bailing_moe import BailingMoeConfig
    from ming_univision_b200.modeling_bailingmm import MingUniVisionForConditionalGeneration

    cfg, vh, tok = synthetic.LLM_WIDE_CONFIG, synthetic.VISHEAD_CONFIG, synthetic.MINGTOK_TINY_CONFIG
    torch.set_default_dtype(BF16)
    try:
        with torch.device(cuda_device):
            m = MingUniVisionForConditionalGeneration(BailingMoeConfig(**cfg), MingTokConfig(**tok), vh)
    finally:
        torch.set_default_dtype(torch.float32)
    sd = {}
    for k, v in synthetic.llm_state_dict(cfg, vh, tok["semantic_decoder"]["embed_dim"], 0).items():
        sd[k if k.startswith("linear_proj.") else "model." + k] = v
    for k, v in synthetic.mingtok_state_dict(tok, 0).items():
        sd["vision." + k] = v
    m.load_state_dict(sd, strict=True)  # copies (and rounds to bf16) tensor by tensor: no second full-size copy
    del sd
    return m


def test_wide_prefill_logits(wide_model, gold, cuda_device):
    """192-token prefill with a 64-token image span (image_gate there), 18 (token, slot) pairs per expert -> the grouped
    tcgen05 expert GEMMs; final-norm hidden states, last-row logits over the full 126464 vocabulary, KV samples."""
    g = gold
    llm = wide_model.model
    ids = torch.from_numpy(g["prefill_ids"]).to(cuda_device)
    S = ids.shape[1]
    rows = torch.from_numpy(g["prefill_rows"]).long()
    cache = llm.new_cache(max_len=256)
    pos = torch.arange(S, device=cuda_device, dtype=torch.int32).unsqueeze(0)
    h = llm.model.forward_tokens(llm.model.embed(ids), pos, cache,
                                 image_mask=torch.from_numpy(g["prefill_image_mask"]).to(cuda_device))
    e_h = rel_l2(h[0].float().cpu()[rows], torch.from_numpy(g["prefill_hidden_rows"]))
    logits = llm.compute_logit(h[:, -1])
    ref = torch.from_numpy(g["prefill_logits_last"])
    e_l = rel_l2(logits[0], ref)
    mad = float((logits[0].float().cpu() - ref).abs().max())
    print(f"wide prefill: hidden rel-L2 {e_h:.3e}, logits rel-L2 {e_l:.3e}, logit max-abs-diff {mad:.3e} "
          f"(ref std {float(ref.std()):.2f}, max {float(ref.abs().max()):.2f}), argmax equal "
          f"{int(logits[0].argmax()) == int(ref.argmax())}")
    assert logits.dtype == torch.float32 and logits.shape[-1] == 126464
    assert e_h < 2e-2 and e_l < 2e-2 and mad < 0.1
    k0 = cache.k[0][0, :, :S].float().cpu()[:, rows]
    v1 = cache.v[1][0, :, :S].float().cpu()[:, rows]
    assert rel_l2(k0, torch.from_numpy(g["prefill_k0_rows"])) < 1e-2
    assert rel_l2(v1, torch.from_numpy(g["prefill_v1_rows"])) < 1.5e-2
    # ---- one cached CFG-row step behind that prefill, B = 2 and B = 3 rows with 2-D padding masks
    D = llm.config.hidden_size
    for B in (2, 3):
        cache.seq_len, cache.batch = S, 1
        cache.repeat_rows(B)
        x1 = torch.randn((1, 1, D), generator=torch.Generator().manual_seed(int(g[f"step{B}_seed"]))).repeat(B, 1, 1)
        mask = torch.from_numpy(g[f"step{B}_mask"]).to(torch.int32).to(cuda_device)
        h2 = llm.model.forward_tokens(x1.to(cuda_device), torch.from_numpy(g[f"step{B}_pos"]).to(cuda_device), cache,
                                      key_mask=mask)
        z = llm.compute_vis_z(h2[:, -1])
        e_h2, e_z = rel_l2(h2, torch.from_numpy(g[f"step{B}_hidden"])), rel_l2(z, torch.from_numpy(g[f"step{B}_z"]))
        print(f"wide CFG step B={B}: hidden rel-L2 {e_h2:.3e}, z rel-L2 {e_z:.3e}")
        assert e_h2 < 2e-2 and e_z < 2e-2
        assert cache.get_seq_length() == S + 1


@pytest.mark.parametrize("grouped", [True, False])
def test_wide_moe_block_operator(wide_model, gold, cuda_device, grouped):
    """The MoE operator boundary `BailingMoeSparseMoeBlock.forward(hidden_states, image_mask)` (reference
    modeling_bailing_moe.py:556-639) at 64 experts / top-6 / I = 1408 + shared 2816, both execution paths: grouped
    tcgen05 GEMMs (the default at 18 pairs per expert) and the weight-streaming kernel."""
    from ming_univision_b200 import ops

    g = gold
    S, D = 192, 2048
    x = torch.randn((1, S, D), generator=torch.Generator().manual_seed(int(g["moe_seed"]))).to(cuda_device).to(BF16)
    im = torch.from_numpy(g["prefill_image_mask"]).to(cuda_device)
    blk = wide_model.model.model.layers[0].mlp
    saved = ops.MOE_GROUPED_MIN_PAIRS_PER_EXPERT
    ops.MOE_GROUPED_MIN_PAIRS_PER_EXPERT = 16 if grouped else 10 ** 9
    try:
        y, (rl, ti) = blk(x, image_mask=im)
    finally:
        ops.MOE_GROUPED_MIN_PAIRS_PER_EXPERT = saved
    rows = torch.from_numpy(g["prefill_rows"]).long()
    e_y = rel_l2(y[0].float().cpu()[rows], torch.from_numpy(g["moe_y_rows"]))
    ref_idx = torch.from_numpy(g["moe_topk_idx"].astype(np.int64))
    same = sum(set(a.tolist()) == set(b.tolist()) for a, b in zip(ti[0].cpu(), ref_idx)) / S
    e_r = rel_l2(rl[0].float().cpu()[rows], torch.from_numpy(g["moe_router_logits_rows"]))
    print(f"wide MoE block ({'grouped' if grouped else 'streaming'}): y rel-L2 {e_y:.3e}, router logits rel-L2 {e_r:.3e}, "
          f"tokens with the reference's expert set {same:.3f}")
    assert ti.dtype == torch.int64 and tuple(ti.shape) == (1, S, 6)
    assert e_r < 1e-2
    assert same >= 0.97   # bf16 router logits may swap the 6th / 7th expert of a token whose margin is below 2^-8
    assert e_y < 2.5e-2


@pytest.mark.parametrize("name", ["t2i", "edit"])
@pytest.mark.parametrize("forced", [True, False])
def test_wide_generate_image(wide_model, gold, cuda_device, name, forced):
    """`generate_image` at the true LLM / RF widths against the reference's own run (B = 2 and B = 3 CFG rows, 4 tokens).

    forced = True: teacher forcing (the reference's latent of step i replaces ours before it is fed back), so every RF
    sample starts from a reference-exact context: latents <= 8e-2, features <= 2e-2, image <= 3e-2.
    forced = False: FREE-RUNNING.  The loop feeds every sampled latent back through the LLM and CFG 3.0 amplifies a
    bf16-sized perturbation from token to token.  The fixture holds the same generation by the reference's own modules
    in ITS bf16 regime (bf16 parameters + bf16 autocast, run on the CPU): its latents leave the fp32 trajectory by
    4.9e-2, 0.10, 0.23, 0.61 (t2i) and 6.8e-2, 0.16, 0.58, 0.86 (edit) over the four tokens.  Ours must stay inside
    that envelope: error(token i) <= 1.5 x the reference-bf16 error + 2e-2."""
    g = gold
    ids = torch.from_numpy(g["gen_ids"]).to(cuda_device)
    noises = [torch.from_numpy(n) for n in g[f"{name}_noises"]]
    ref_l, ref_f = torch.from_numpy(g[f"{name}_latents"]), torch.from_numpy(g[f"{name}_feats"])
    bf_l = torch.from_numpy(g[f"{name}_latents_bf16"])
    lats, feats = [], []
    vision = wide_model.vision
    orig = vision.forward_feature_decoder

    def spy(latent, past_key_values=None):
        i = len(lats)
        lats.append(latent.float().cpu())
        if forced:
            latent = ref_l[:, i:i + 1].to(cuda_device)
        r = orig(latent, past_key_values=past_key_values)
        feats.append(r["x_norm_patchtokens"].float().cpu())
        return r

    vision.forward_feature_decoder = spy
    try:
        img, fmask = wide_model.generate_image_from_prompt(
            ids, uncond_attention_mask=torch.from_numpy(g[f"{name}_uncond"]).to(cuda_device),
            text_uncond_attention_mask=torch.from_numpy(g[f"{name}_text_uncond"]).to(cuda_device),
            image_gen_temperature=0.9, noises=noises)
    finally:
        vision.forward_feature_decoder = orig
    B, n_tok = ref_l.shape[0], ref_l.shape[1]
    assert len(lats) == n_tok and fmask.shape[0] == B
    assert torch.equal(fmask.cpu().long(), torch.from_numpy(g[f"{name}_final_mask"]))
    assert wide_model.past_key_values.get_seq_length() == int(g[f"{name}_cache_len"])
    e_l = [rel_l2(lats[i], ref_l[:, i:i + 1]) for i in range(n_tok)]
    e_f = [rel_l2(feats[i], ref_f[:, i:i + 1]) for i in range(n_tok)]
    e_bf = [rel_l2(bf_l[:, i:i + 1], ref_l[:, i:i + 1]) for i in range(n_tok)]
    print(f"wide generate_image {name} B={B} {'forced' if forced else 'free'}: latent err/token "
          f"{['%.2e' % e for e in e_l]} (reference bf16 regime: {['%.2e' % e for e in e_bf]}) feat err/token "
          f"{['%.2e' % e for e in e_f]}")
    if forced:
        assert max(e_l) < 8e-2, e_l
        assert max(e_f) < 2e-2, e_f
        assert rel_l2(img, torch.from_numpy(g[f"{name}_image"])[0:1]) < 3e-2
    else:
        assert all(e <= 1.5 * b + 2e-2 for e, b in zip(e_l, e_bf)), (e_l, e_bf)
        assert bool(torch.isfinite(img.float()).all())


def test_wide_cfg_rows_computed_once_is_bit_identical(wide_model, gold, cuda_device):
    """The CFG rows carry identical latents, so the semantic-decoder step, linear_proj and the pixel decoder run on ONE
    row (BailingMoeForCausalLM.dedupe_cfg_rows) — this must reproduce the B-row computation bit for bit, on the eager
    loop and on the CUDA-graph fast path."""
    g = gold
    ids = torch.from_numpy(g["gen_ids"]).to(cuda_device)
    noises = [torch.from_numpy(n) for n in g["edit_noises"]]
    um = torch.from_numpy(g["edit_uncond"]).to(cuda_device)
    tm = torch.from_numpy(g["edit_text_uncond"]).to(cuda_device)
    llm = wide_model.model
    outs = {}
    try:
        for graph in (False, True):
            for dedupe in (False, True):
                llm.use_cuda_graph, llm.dedupe_cfg_rows = graph, dedupe
                img, fmask = wide_model.generate_image_from_prompt(ids, uncond_attention_mask=um,
                                                                   text_uncond_attention_mask=tm,
                                                                   image_gen_temperature=0.9, noises=noises)
                T = wide_model.past_key_values.get_seq_length()
                outs[(graph, dedupe)] = (img.float().cpu(), fmask.cpu(),
                                         wide_model.past_key_values.k[1][0, :, :T].float().cpu())
    finally:
        llm.use_cuda_graph, llm.dedupe_cfg_rows = True, True
    base = outs[(False, False)]
    for key, o in outs.items():
        assert torch.equal(o[0], base[0]), key
        assert torch.equal(o[1], base[1]), key
        assert torch.equal(o[2], base[2]), key


def test_state_dict_reload_drops_packs_and_graphs(wide_model, gold, cuda_device):
    """loading a checkpoint through the PARENT wrapper after a forward / graph capture must not keep serving
    the old packed expert slabs, RF packs, MingTok packs or captured graphs."""
    g = gold
    ids = torch.from_numpy(g["gen_ids"]).to(cuda_device)
    noises = [torch.from_numpy(n) for n in g["t2i_noises"]]
    um = torch.from_numpy(g["t2i_uncond"]).to(cuda_device)
    tm = torch.from_numpy(g["t2i_text_uncond"]).to(cuda_device)

    def run():
        img, _ = wide_model.generate_image_from_prompt(ids, uncond_attention_mask=um, text_uncond_attention_mask=tm,
                                                       image_gen_temperature=0.9, noises=noises)
        return img.float().cpu()

    a = run()
    sd = {k: v.clone() for k, v in wide_model.state_dict().items()}
    # one tensor of every pack owner at a time: routed-expert slab, RF pack, MingTok pack, linear_proj pack
    for k in ("model.diffloss.net.res_blocks.0.mlp.w3.weight", "vision.semantic_decoder.blocks.0.0.mlp.w3.weight",
              "linear_proj.2.weight", "model.model.layers.1.mlp.shared_experts.down_proj.weight"):
        mod = dict(sd)
        mod[k] = torch.zeros_like(sd[k])
        try:
            wide_model.load_state_dict(mod, strict=True)
            b = run()
            assert not torch.equal(a, b), f"{k}: the modified weights were ignored (stale packs / graphs)"
        finally:
            wide_model.load_state_dict(sd, strict=True)
    # all 64 routed experts of layer 1 zeroed -> whatever the router picks, the output must change
    mod = dict(sd)
    for k in sd:
        if k.startswith("model.model.layers.1.mlp.experts.") and k.endswith("down_proj.weight"):
            mod[k] = torch.zeros_like(sd[k])
    try:
        wide_model.load_state_dict(mod, strict=True)
        assert not torch.equal(a, run()), "routed-expert slabs were not re-packed"
    finally:
        wide_model.load_state_dict(sd, strict=True)
    assert torch.equal(a, run())


@pytest.mark.parametrize("name,G", [("edit", 2), ("t2i", 3)])
def test_wide_batched_generation_equals_single(wide_model, gold, cuda_device, name, G):
    """SURVEY.md §8f.1 (the reference asserts one sequence, modeling_bailing_moe.py:1865): G requests generated TOGETHER
    (G x CFG rows = 6 rows per pass over the LLM / RF-head / semantic-decoder weights) must give every request exactly the
    image it gets on its own — all kernels of the step are row-independent, so this is bit-for-bit."""
    g = gold
    base = torch.from_numpy(g["gen_ids"])
    gen = torch.Generator().manual_seed(77)
    ids = torch.cat([base] + [torch.randint(0, 100000, base.shape, generator=gen) for _ in range(G - 1)]).to(cuda_device)
    n_tok = wide_model.model.config.num_image_tokens_for_gen
    noises = [torch.randn((G, 32), generator=gen) for _ in range(n_tok + 1)]
    um = torch.from_numpy(g[f"{name}_uncond"]).to(cuda_device)
    tm = torch.from_numpy(g[f"{name}_text_uncond"]).to(cuda_device)
    singles = []
    for i in range(G):
        img, fmask = wide_model.generate_image_from_prompt(ids[i:i + 1], uncond_attention_mask=um,
                                                           text_uncond_attention_mask=tm, image_gen_temperature=0.9,
                                                           noises=[n[i:i + 1] for n in noises])
        singles.append((img.float().cpu(), fmask.cpu()))
    for graph in (True, False):
        wide_model.model.use_cuda_graph = graph
        try:
            imgs, fmask = wide_model.generate_image_from_prompt(ids, uncond_attention_mask=um.expand(G, -1),
                                                                text_uncond_attention_mask=tm.expand(G, -1),
                                                                image_gen_temperature=0.9, noises=noises)
        finally:
            wide_model.model.use_cuda_graph = True
        B = singles[0][1].shape[0]
        assert tuple(imgs.shape) == (G,) + tuple(singles[0][0].shape[1:]) and fmask.shape[0] == G * B
        assert wide_model.past_key_values.batch == G
        for i in range(G):
            assert torch.equal(imgs[i:i + 1].float().cpu(), singles[i][0]), (graph, i)
            assert torch.equal(fmask[i * B:(i + 1) * B].cpu(), singles[i][1]), (graph, i)


def test_rf_sampler_groups_equal_single_samples(cuda_device):
    """RectifiedFlowLoss.sample(groups=G): G samples of B CFG rows through ONE pass over the weights == G separate calls."""
    from ming_univision_b200.diff_loss_rf_swiglu import RectifiedFlowLoss

    cfg = synthetic.RF_CONFIG
    with torch.device(cuda_device):
        m = RectifiedFlowLoss(cfg["target_channels"], cfg["z_channels"], cfg["depth"], cfg["width"],
                              str(cfg["num_sampling_steps"]), mlp_mult=cfg["mlp_mult"])
    m.load_state_dict({k: v.to(cuda_device) for k, v in synthetic.rf_state_dict(cfg, 0).items()})
    m = m.to(BF16)
    gen = torch.Generator().manual_seed(3)
    for B, G in ((3, 2), (2, 3), (1, 4)):
        z = torch.randn((G * B, cfg["z_channels"]), generator=gen).to(cuda_device)
        noise = torch.randn((G, 32), generator=gen).to(cuda_device)
        tc = 3.0 if B > 1 else 1.0
        both = m.sample(z, temperature=0.9, text_cfg=tc, image_cfg=1.1, groups=G,
                        noise=noise if B > 1 else noise.repeat_interleave(B, 0))
        for i in range(G):
            one = m.sample(z[i * B:(i + 1) * B], temperature=0.9, text_cfg=tc, image_cfg=1.1, noise=noise[i:i + 1])
            assert torch.equal(both[i * B:(i + 1) * B], one), (B, G, i)
