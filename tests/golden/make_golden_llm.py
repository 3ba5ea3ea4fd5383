"""Golden fixtures for the Bailing-MoE AR path from the UNMODIFIED reference (run in the build container):
    python tests/golden/make_golden_llm.py   ->   tests/golden/llm_tiny.npz

A 2-layer, 128-wide BailingMoeForCausalLM (8 experts top-2 + 1 shared, GQA 4q/2kv x 32, multi_gate) with vis_head,
a tiny RF head, the tiny MingTok and a linear_proj is driven through the reference's own code:
  * prefill of a 10-token prompt (with an image-token span routed by image_gate) -> hidden states, logits, KV cache
  * one cached decode step with B = 2 rows and a 2-D padding mask (the CFG-row situation)
  * the reference's `generate_image` end to end (4 visual tokens, RF sampler, MingTok callbacks) for the B = 2 (T2I)
    and B = 3 (edit) mask configurations: per-token latents and features, final image, final mask, cache length.
The torch.randn draws inside RectifiedFlowLoss.sample are reproduced from the same seed and stored as `noises`.
"""
import contextlib
import io
import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")

from ming_univision_b200 import synthetic  # noqa: E402
from oracle import ref_shims  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    torch.set_num_threads(8)
    cfg, vh = synthetic.LLM_TINY_CONFIG, synthetic.VISHEAD_TINY_CONFIG
    tok_cfg = synthetic.MINGTOK_TINY_CONFIG
    F_dim = tok_cfg["semantic_decoder"]["embed_dim"]
    sd = synthetic.llm_state_dict(cfg, vh, feature_dim=F_dim, seed=0)
    llm_sd = {k: v for k, v in sd.items() if not k.startswith("linear_proj.")}
    model, Legacy = ref_shims.build_reference_llm(cfg, vh, None)
    missing = model.load_state_dict(llm_sd, strict=False)
    assert all("rotary_emb" in k for k in missing.missing_keys) and not missing.unexpected_keys, missing
    mingtok = ref_shims.build_reference_mingtok(tok_cfg, synthetic.mingtok_state_dict(tok_cfg, 0), fa_enable=False)
    D = cfg["hidden_size"]
    lin = torch.nn.Sequential(torch.nn.Linear(F_dim, D), torch.nn.GELU(), torch.nn.Linear(D, D)).eval()
    lin.load_state_dict({"0.weight": sd["linear_proj.0.weight"], "0.bias": sd["linear_proj.0.bias"],
                         "2.weight": sd["linear_proj.2.weight"], "2.bias": sd["linear_proj.2.bias"]})
    out = {}
    g = torch.Generator().manual_seed(21)

    # ---- prefill (image_mask over tokens 3..6 -> image_gate there)
    S = 10
    ids = torch.randint(0, 400, (1, S), generator=g)
    emb = model.model.word_embeddings(ids)
    image_mask = torch.zeros((1, S), dtype=torch.bool)
    image_mask[:, 3:7] = True
    with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
        o = model.model(inputs_embeds=emb, attention_mask=torch.ones(1, S, dtype=torch.long), use_cache=True,
                        past_key_values=Legacy(), image_mask=image_mask)
        logits = model.compute_logit(o.last_hidden_state)
    out["prefill_ids"], out["prefill_image_mask"] = ids.numpy(), image_mask.numpy()
    out["prefill_hidden"], out["prefill_logits_last"] = o.last_hidden_state.numpy(), logits[:, -1].float().numpy()
    out["prefill_k0"] = o.past_key_values.key_cache[0].numpy()
    out["prefill_v1"] = o.past_key_values.value_cache[1].numpy()

    # ---- one decode step with CFG rows: B = 2, row 1 masks part of the prompt
    cache = Legacy.from_legacy_cache(tuple((k.repeat(2, 1, 1, 1), v.repeat(2, 1, 1, 1))
                                           for k, v in o.past_key_values.to_legacy_cache()))
    mask = torch.ones((2, S + 1), dtype=torch.long)
    mask[1, 2:8] = 0
    x1 = torch.randn((1, 1, D), generator=g).repeat(2, 1, 1)
    pos = (mask.cumsum(-1) - 1)[:, -1:]
    with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
        o2 = model.model(inputs_embeds=x1, attention_mask=mask, position_ids=pos, use_cache=True, past_key_values=cache)
        z = model.vis_head(o2.last_hidden_state[:, -1:])
    out["step_x"], out["step_mask"], out["step_pos"] = x1.numpy(), mask.numpy(), pos.numpy()
    out["step_hidden"], out["step_z"] = o2.last_hidden_state.numpy(), z.numpy()

    # ---- generate_image end to end (B = 2 and B = 3)
    n_tok = cfg["num_image_tokens_for_gen"]
    for name, uncond, text_uncond in (("t2i", [1, 1, 0, 0, 0, 0, 0, 0, 1, 1, 1], None),
                                      ("edit", [1, 1, 0, 0, 0, 0, 0, 0, 1, 1, 1], [1, 1, 1, 1, 1, 0, 0, 0, 1, 1, 1])):
        with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
            pre = model.model(inputs_embeds=emb, attention_mask=torch.ones(1, S, dtype=torch.long), use_cache=True,
                              past_key_values=Legacy())
        start = model.model.word_embeddings(torch.tensor([[cfg["image_start_token"]]]))
        am = torch.ones((1, S + 1), dtype=torch.long)
        um = torch.tensor([uncond], dtype=torch.long)
        tm = torch.tensor([text_uncond], dtype=torch.long) if text_uncond is not None else torch.zeros_like(um)
        lats, feats = [], []

        def l2s(latent, past_key_values=None):
            r = mingtok.forward_feature_decoder(latent, past_key_values=past_key_values)
            lats.append(latent.clone())
            feats.append(r["x_norm_patchtokens"].clone())
            return r

        torch.manual_seed(11)
        noises = torch.stack([torch.randn(1, 32) for _ in range(n_tok + 1)])
        torch.manual_seed(11)
        with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
            img, mo, fmask = model.generate_image(
                input_embeds=start, past_key_values=pre.past_key_values, attention_mask=am, uncond_attention_mask=um,
                text_uncond_attention_mask=tm, latent_to_sem_func=l2s, linear_proj=lin,
                sem_to_pix_func=mingtok.forward_pixel_decoder, image_gen_temperature=0.9)
        out[f"{name}_uncond"], out[f"{name}_text_uncond"] = um.numpy(), tm.numpy()
        out[f"{name}_noises"] = noises.numpy()
        out[f"{name}_latents"] = torch.cat(lats, dim=1).numpy()
        out[f"{name}_feats"] = torch.cat(feats, dim=1).numpy()
        out[f"{name}_image"] = img.numpy()
        out[f"{name}_final_mask"] = fmask.numpy()
        out[f"{name}_last_hidden"] = mo.last_hidden_state.numpy()
        out[f"{name}_cache_len"] = np.array(mo.past_key_values.get_seq_length())
        out[f"{name}_cache_batch"] = np.array(mo.past_key_values.key_cache[0].shape[0])
        print(name, "rows", fmask.shape[0], "latents", out[f"{name}_latents"].shape, "image", tuple(img.shape),
              "cache", int(out[f"{name}_cache_len"]), "batch", int(out[f"{name}_cache_batch"]))
    np.savez_compressed(os.path.join(OUT, "llm_tiny.npz"), seed=0, **out)


if __name__ == "__main__":
    main()
