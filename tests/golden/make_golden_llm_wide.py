"""Golden fixtures for the Bailing-MoE AR path at the TRUE 16B-A3B widths, depth-reduced, from the
UNMODIFIED reference run on the CPU of the build container:

    python tests/golden/make_golden_llm_wide.py   ->   tests/golden/llm_wide.npz        (~2 min, ~25 GB of host memory)

Model: `synthetic.LLM_WIDE_CONFIG` = mingunivision/config.json's llm_config with 2 layers instead of 28 — hidden 2048,
16 query / 4 KV heads x 128, 64 routed experts top-6 (I = 1408) + shared expert 2816, multi_gate, vocabulary 126464 —
plus the default-size vis_head (2048 -> 3072 + LayerNorm) and RF head (1.285 B parameters, 16 Euler steps), the tiny
MingTok (feature width 128) and a linear_proj 128 -> 2048.  Driven through the reference's own modules
(modeling_bailing_moe.py:556-639, 1165-1239, 1391-1540, 1604-1673, 1844-1965):

  * prefill of a 192-token prompt with a 64-token image span (image_gate there; 18 pairs per expert -> our grouped
    tcgen05 expert GEMMs): final-norm hidden states (every 8th row + the last), last-row logits (all 126464), KV samples
  * `BailingMoeSparseMoeBlock.forward` of layer 0 on a seeded [1, 192, 2048] input: output, router top-k ids
  * one cached CFG-row step with B = 2 and with B = 3 rows (2-D padding masks): hidden, z = vis_head(h)
  * `generate_image` end to end, 4 visual tokens, B = 2 (t2i) and B = 3 (edit): per-token latents / features, image,
    mask, cache bookkeeping
  * the same two generations by the SAME reference modules in bf16 (parameters cast to bf16, CPU bf16 autocast — the
    reference's "R2" GPU regime, SURVEY.md Appendix C, emulated on the CPU) free-running: `{name}_latents_bf16`.  Its
    per-token distance from the fp32 run is the drift envelope the free-running GPU test is held to.
Large tensors are stored as strided samples so the file stays small; the test regenerates inputs from the seeds.
"""
import contextlib
import copy
import io
import os
import sys
import time
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")

from ming_univision_b200 import synthetic  # noqa: E402
from oracle import ref_shims  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
S = 192
IMG_SPAN = (40, 104)
ROW_STRIDE = 8


def quiet():
    return contextlib.redirect_stdout(io.StringIO())


def main(tag="wide", cfg=None, vh=None):
    torch.set_num_threads(max(1, os.cpu_count() or 8))
    cfg = cfg or synthetic.LLM_WIDE_CONFIG
    vh = vh or synthetic.VISHEAD_CONFIG
    tok_cfg = synthetic.MINGTOK_TINY_CONFIG
    F_dim = tok_cfg["semantic_decoder"]["embed_dim"]
    t0 = time.time()
    sd = synthetic.llm_state_dict(cfg, vh, feature_dim=F_dim, seed=0)
    print(f"weights: {sum(v.numel() for v in sd.values()) / 1e9:.2f} B parameters in {time.time() - t0:.0f} s", flush=True)
    llm_sd = {k: v for k, v in sd.items() if not k.startswith("linear_proj.")}
    model, Legacy = ref_shims.build_reference_llm(cfg, vh, None)
    missing = model.load_state_dict(llm_sd, strict=False)
    assert all("rotary_emb" in k for k in missing.missing_keys) and not missing.unexpected_keys, missing
    mingtok = ref_shims.build_reference_mingtok(tok_cfg, synthetic.mingtok_state_dict(tok_cfg, 0), fa_enable=False)
    D = cfg["hidden_size"]
    lin = torch.nn.Sequential(torch.nn.Linear(F_dim, D), torch.nn.GELU(), torch.nn.Linear(D, D)).eval()
    lin.load_state_dict({"0.weight": sd["linear_proj.0.weight"], "0.bias": sd["linear_proj.0.bias"],
                         "2.weight": sd["linear_proj.2.weight"], "2.bias": sd["linear_proj.2.bias"]})
    del sd, llm_sd
    out = {}
    g = torch.Generator().manual_seed(21)

    # ---- prefill
    ids = torch.randint(0, 100000, (1, S), generator=g)
    emb = model.model.word_embeddings(ids)
    image_mask = torch.zeros((1, S), dtype=torch.bool)
    image_mask[:, IMG_SPAN[0]:IMG_SPAN[1]] = True
    with torch.no_grad(), quiet():
        o = model.model(inputs_embeds=emb, attention_mask=torch.ones(1, S, dtype=torch.long), use_cache=True,
                        past_key_values=Legacy(), image_mask=image_mask)
        logits = model.compute_logit(o.last_hidden_state[:, -1:])
    rows = sorted(set(range(0, S, ROW_STRIDE)) | {S - 1})
    out["prefill_ids"], out["prefill_image_mask"] = ids.numpy(), image_mask.numpy()
    out["prefill_rows"] = np.array(rows)
    out["prefill_hidden_rows"] = o.last_hidden_state[0, rows].numpy()
    out["prefill_logits_last"] = logits[0, -1].float().numpy()
    out["prefill_k0_rows"] = o.past_key_values.key_cache[0][0, :, rows].numpy()
    out["prefill_v1_rows"] = o.past_key_values.value_cache[1][0, :, rows].numpy()
    print(f"prefill done ({time.time() - t0:.0f} s); logits std {float(logits.std()):.3f} max {float(logits.abs().max()):.2f}",
          flush=True)

    # ---- the MoE operator boundary: BailingMoeSparseMoeBlock.forward(hidden_states, image_mask) of layer 0
    xm = torch.randn((1, S, D), generator=torch.Generator().manual_seed(33))
    with torch.no_grad(), quiet():
        ym, (rl, ti) = model.model.layers[0].mlp(xm, image_mask=image_mask)
    out["moe_seed"] = np.array(33)
    out["moe_y_rows"] = ym[0, rows].numpy()
    out["moe_topk_idx"] = ti.reshape(S, -1).numpy().astype(np.int16)
    out["moe_router_logits_rows"] = rl.reshape(S, -1)[rows].float().numpy()

    # ---- one cached CFG-row decode step: B = 2 and B = 3 rows with 2-D padding masks
    for B in (2, 3):
        cache = Legacy.from_legacy_cache(tuple((k.repeat(B, 1, 1, 1), v.repeat(B, 1, 1, 1))
                                               for k, v in o.past_key_values.to_legacy_cache()))
        mask = torch.ones((B, S + 1), dtype=torch.long)
        mask[1, 2:S - 3] = 0
        if B == 3:
            mask[2, 2:IMG_SPAN[0]] = 0
        x1 = torch.randn((1, 1, D), generator=torch.Generator().manual_seed(40 + B)).repeat(B, 1, 1)
        pos = (mask.cumsum(-1) - 1)[:, -1:]
        with torch.no_grad(), quiet():
            o2 = model.model(inputs_embeds=x1, attention_mask=mask, position_ids=pos, use_cache=True,
                             past_key_values=cache)
            z = model.vis_head(o2.last_hidden_state[:, -1:])
        out[f"step{B}_seed"] = np.array(40 + B)
        out[f"step{B}_mask"], out[f"step{B}_pos"] = mask.numpy(), pos.numpy()
        out[f"step{B}_hidden"], out[f"step{B}_z"] = o2.last_hidden_state.numpy(), z.reshape(B, -1).numpy()
    print(f"CFG steps done ({time.time() - t0:.0f} s)", flush=True)

    # ---- generate_image end to end (B = 2 and B = 3), fp32 and the bf16 regime of the same reference modules
    n_tok = cfg["num_image_tokens_for_gen"]
    P = 24  # prompt length of the generation cases
    ids_g = ids[:, :P]
    uncond = [1, 1] + [0] * (P - 5) + [1, 1, 1, 1]           # P + 1 entries (prompt + <image> start token)
    text_uncond = [1] * 8 + [0] * (P - 11) + [1, 1, 1, 1]
    out["gen_ids"] = ids_g.numpy()
    torch.manual_seed(11)
    noises = torch.stack([torch.randn(1, 32) for _ in range(n_tok + 1)])
    model_bf, mingtok_bf, lin_bf = None, None, None
    for name, tu in (("t2i", None), ("edit", text_uncond)):
        um = torch.tensor([uncond], dtype=torch.long)
        tm = torch.tensor([tu], dtype=torch.long) if tu is not None else torch.zeros_like(um)
        out[f"{name}_uncond"], out[f"{name}_text_uncond"] = um.numpy(), tm.numpy()
        out[f"{name}_noises"] = noises.numpy()
        for regime in ("fp32", "bf16"):
            if regime == "fp32":
                mdl, tokm, lp, ctx, dt = model, mingtok, lin, contextlib.nullcontext(), torch.float32
            else:
                if model_bf is None:
                    model_bf = copy.deepcopy(model).to(torch.bfloat16)
                    # MingTok keeps fp32 parameters: CPU layer_norm refuses bf16 weights with the fp32 rows the RF sampler
                    # hands over; under autocast its linears still run in bf16 and its LayerNorms in fp32 (= CUDA autocast)
                    mingtok_bf = mingtok
                    lin_bf = copy.deepcopy(lin).to(torch.bfloat16)
                mdl, tokm, lp, dt = model_bf, mingtok_bf, lin_bf, torch.bfloat16
                ctx = torch.autocast("cpu", dtype=torch.bfloat16)
            lats, feats = [], []

            def l2s(latent, past_key_values=None, tokm=tokm, lats=lats, feats=feats):
                r = tokm.forward_feature_decoder(latent, past_key_values=past_key_values)
                lats.append(latent.float().clone())
                feats.append(r["x_norm_patchtokens"].float().clone())
                return r

            with torch.no_grad(), quiet(), contextlib.redirect_stderr(io.StringIO()), ctx:
                pre = mdl.model(inputs_embeds=mdl.model.word_embeddings(ids_g),
                                attention_mask=torch.ones(1, P, dtype=torch.long), use_cache=True,
                                past_key_values=Legacy())
                start = mdl.model.word_embeddings(torch.tensor([[cfg["image_start_token"]]]))
                torch.manual_seed(11)
                img, mo, fmask = mdl.generate_image(
                    input_embeds=start, past_key_values=pre.past_key_values,
                    attention_mask=torch.ones((1, P + 1), dtype=torch.long), uncond_attention_mask=um,
                    text_uncond_attention_mask=tm, latent_to_sem_func=l2s, linear_proj=lp,
                    sem_to_pix_func=tokm.forward_pixel_decoder, image_gen_temperature=0.9)
            sfx = "" if regime == "fp32" else "_bf16"
            out[f"{name}_latents{sfx}"] = torch.cat(lats, dim=1).numpy()
            out[f"{name}_feats{sfx}"] = torch.cat(feats, dim=1).numpy()
            out[f"{name}_image{sfx}"] = img.float().numpy()
            if regime == "fp32":
                out[f"{name}_final_mask"] = fmask.numpy()
                out[f"{name}_last_hidden"] = mo.last_hidden_state.float().numpy()
                out[f"{name}_cache_len"] = np.array(mo.past_key_values.get_seq_length())
            else:
                a, b = out[f"{name}_latents"], out[f"{name}_latents_bf16"]
                drift = [float(np.linalg.norm(a[:, i] - b[:, i]) / np.linalg.norm(a[:, i])) for i in range(a.shape[1])]
                print(f"{name}: bf16-regime reference vs fp32 reference, latent rel-L2 per token: "
                      f"{['%.2e' % d for d in drift]}", flush=True)
        print(f"{name} done ({time.time() - t0:.0f} s): rows {fmask.shape[0]} latents {out[name + '_latents'].shape} "
              f"image {tuple(img.shape)}", flush=True)
    path = os.path.join(OUT, f"llm_{tag}.npz")
    np.savez_compressed(path, seed=0, **out)
    print(f"wrote {path}: {os.path.getsize(path) / 1e6:.2f} MB", flush=True)


if __name__ == "__main__":
    main()
