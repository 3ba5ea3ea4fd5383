"""Per-token error of the product generate_image against the golden reference trajectory (tiny model), free-running
and teacher-forced (the reference's latent of step i replaces ours before it is fed back)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from ming_univision_b200 import synthetic
from ming_univision_b200.mingtok import MingTokConfig
from ming_univision_b200.modeling_bailing_moe import BailingMoeConfig
from ming_univision_b200.modeling_bailingmm import MingUniVisionForConditionalGeneration
from parity_metrics import rel_l2

dev = torch.device("cuda:0")
cfg, vh, tok = synthetic.LLM_TINY_CONFIG, synthetic.VISHEAD_TINY_CONFIG, synthetic.MINGTOK_TINY_CONFIG
with torch.device(dev):
    m = MingUniVisionForConditionalGeneration(BailingMoeConfig(**cfg), MingTokConfig(**tok), vh)
sd = {}
for k, v in synthetic.llm_state_dict(cfg, vh, tok["semantic_decoder"]["embed_dim"], 0).items():
    sd[k if k.startswith("linear_proj.") else "model." + k] = v
for k, v in synthetic.mingtok_state_dict(tok, 0).items():
    sd["vision." + k] = v
m.load_state_dict({k: v.to(dev) for k, v in sd.items()}, strict=True)
m = m.to(torch.bfloat16)
g = np.load("tests/golden/llm_tiny.npz")
ids = torch.from_numpy(g["prefill_ids"]).to(dev)
for name in ("t2i", "edit"):
    for forced in (False, True):
        ref_l = torch.from_numpy(g[f"{name}_latents"])
        ref_f = torch.from_numpy(g[f"{name}_feats"])
        noises = [torch.from_numpy(n) for n in g[f"{name}_noises"]]
        lats, feats = [], []
        orig = m.vision.forward_feature_decoder
        def spy(latent, past_key_values=None):
            i = len(lats)
            lats.append(latent.float().cpu())
            if forced:
                latent = ref_l[:, i:i + 1].to(dev)
            r = orig(latent, past_key_values=past_key_values)
            feats.append(r["x_norm_patchtokens"].float().cpu())
            return r
        m.vision.forward_feature_decoder = spy
        try:
            m.generate_image_from_prompt(ids, uncond_attention_mask=torch.from_numpy(g[f"{name}_uncond"]).to(dev),
                                         text_uncond_attention_mask=torch.from_numpy(g[f"{name}_text_uncond"]).to(dev),
                                         image_gen_temperature=0.9, noises=noises)
        finally:
            m.vision.forward_feature_decoder = orig
        el = [rel_l2(lats[i], ref_l[:, i:i + 1]) for i in range(len(lats))]
        ef = [rel_l2(feats[i], ref_f[:, i:i + 1]) for i in range(len(feats))]
        print(name, "forced" if forced else "free  ", "latent err/token", ["%.3e" % e for e in el], "feat err/token", ["%.3e" % e for e in ef], flush=True)
