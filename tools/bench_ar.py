"""Secondary measurement (BASELINE configs[3] shape on ONE GPU, experts not sharded): text->image AR generation of 256
continuous visual tokens with the full-size Bailing-MoE 16B-A3B (random bf16 weights generated on the device), the
default RF head (16 Euler steps), the MingTok cached semantic decoder and the pixel decoder, B = 2 CFG rows.
Reports visual tokens/s and a per-stage breakdown.  Development / documentation tool — the judged line is bench.py."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ming_univision_b200 import _lib, synthetic  # noqa: E402
from ming_univision_b200.mingtok import MingTokConfig  # noqa: E402
from ming_univision_b200.modeling_bailing_moe import BailingMoeConfig  # noqa: E402
from ming_univision_b200.modeling_bailingmm import MingUniVisionForConditionalGeneration  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    layers = int(os.environ.get("AR_LAYERS", "28"))
    n_tok = int(os.environ.get("AR_TOKENS", "256"))
    llm_cfg = dict(synthetic.LLM_CONFIG, num_hidden_layers=layers, num_image_tokens_for_gen=n_tok)
    t0 = time.time()
    torch.set_default_dtype(torch.bfloat16)
    with torch.device(dev):
        m = MingUniVisionForConditionalGeneration(BailingMoeConfig(**llm_cfg), MingTokConfig(**synthetic.MINGTOK_CONFIG),
                                                  synthetic.VISHEAD_CONFIG)
    torch.set_default_dtype(torch.float32)
    g = torch.Generator(device=dev).manual_seed(0)
    with torch.no_grad():
        for name, p in m.named_parameters():  # O(1) activations: N(0, 1/fan_in) linears, unit norms
            if p.dim() >= 2:
                p.copy_(torch.randn(p.shape, generator=g, device=dev, dtype=torch.float32) / p.shape[-1] ** 0.5)
            elif "norm" in name or "ln" in name:
                p.fill_(1.0 if name.endswith("weight") else 0.0)
            else:
                p.zero_()
    torch.cuda.synchronize()
    print(f"model built in {time.time() - t0:.1f} s, {torch.cuda.memory_allocated() / 1e9:.1f} GB", flush=True)
    S = 40
    ids = torch.randint(0, 100000, (1, S), device=dev)
    um = torch.ones((1, S + 1), dtype=torch.int32, device=dev)
    um[:, 2:S - 2] = 0

    def run():
        return m.generate_image_from_prompt(ids, uncond_attention_mask=um, text_uncond_attention_mask=torch.zeros_like(um),
                                            image_gen_temperature=1.0)

    run()  # warm-up: packs weights, captures the RF CUDA graph
    torch.cuda.synchronize()
    l0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    img, fmask = run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    res = {"workload": f"T2I AR generation, {n_tok} visual tokens, B=2 CFG rows, {layers}-layer Bailing-MoE, 1 GPU",
           "ms_per_image": round(ms, 1), "visual_tokens_per_s": round(n_tok / (ms / 1e3), 2),
           "ms_per_token": round(ms / n_tok, 3), "kernel_launches": _lib.launch_count() - l0,
           "image_shape": list(img.shape), "mem_gb": round(torch.cuda.max_memory_allocated() / 1e9, 1)}
    print(json.dumps(res), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/bench_ar.json", "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
