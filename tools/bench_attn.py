"""Attention A/B on one B200: tcgen05 / TMEM kernel (backend 1) vs the warp-level mma.sync kernel (backend 2) vs
torch SDPA (flash-attn-2 class library kernel), on the MingTok shapes of BASELINE configs[1] and the LLM prefill shape.
Development / documentation tool."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ming_univision_b200 import ops  # noqa: E402


def timeit(fn, iters=30, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def main():
    dev = torch.device("cuda:0")
    res = []
    for name, B, S, H, causal in [("enc.attn", 64, 65, 12, False), ("sem.attn", 64, 65, 16, True),
                                  ("pix.attn", 64, 256, 16, False), ("enc512.attn", 16, 257, 12, False),
                                  ("pix512.attn", 16, 1024, 16, False)]:
        if os.environ.get("ATTN_ONLY") and os.environ["ATTN_ONLY"] != name:
            continue
        qkv = torch.randn((B, S, 3 * H * 64), device=dev).to(torch.bfloat16)
        fl = 4.0 * B * H * S * S * 64 * (0.5 if causal else 1.0)
        row = {"op": name, "B": B, "S": S, "H": H, "causal": causal}
        outs = {}
        for bk, key in ((1, "tcgen05"), (2, "mma_sync")):
            ops.set_attn_backend(bk)
            outs[key] = ops.attention_hd64(qkv, B, S, H, causal)
            ms = timeit(lambda: ops.attention_hd64(qkv, B, S, H, causal))
            row[key + "_ms"] = round(ms, 4)
            row[key + "_tflops"] = round(fl / ms / 1e9, 1)
        ops.set_attn_backend(0)
        row["max_abs_diff"] = float((outs["tcgen05"].float() - outs["mma_sync"].float()).abs().max())
        q, k, v = qkv.view(B, S, 3, H, 64).permute(2, 0, 3, 1, 4)
        row["sdpa_ms"] = round(timeit(lambda: torch.nn.functional.scaled_dot_product_attention(q, k, v, is_causal=causal)), 4)
        res.append(row)
        print(row, flush=True)
    if os.environ.get("ATTN_ONLY"):
        return
    # LLM prefill: GQA 16 q / 4 kv heads, head_dim 128, causal, S = 1552
    B, S, H, Hkv, hd = 1, 1552, 16, 4, 128
    q = torch.randn((B * S, H * hd), device=dev).to(torch.bfloat16)
    kc = torch.randn((B, Hkv, S + 64, hd), device=dev).to(torch.bfloat16)
    vc = torch.randn((B, Hkv, S + 64, hd), device=dev).to(torch.bfloat16)
    fl = 4.0 * B * H * S * S * hd * 0.5
    row = {"op": "llm.prefill.attn", "B": B, "S": S, "H": H, "Hkv": Hkv, "hd": hd}
    outs = {}
    for bk, key in ((1, "tcgen05"), (2, "mma_sync")):
        ops.set_attn_backend(bk)
        outs[key] = ops.attn_prefill_gqa(q, kc, vc, B, S, H)
        ms = timeit(lambda: ops.attn_prefill_gqa(q, kc, vc, B, S, H))
        row[key + "_ms"] = round(ms, 4)
        row[key + "_tflops"] = round(fl / ms / 1e9, 1)
    ops.set_attn_backend(0)
    row["max_abs_diff"] = float((outs["tcgen05"].float() - outs["mma_sync"].float()).abs().max())
    res.append(row)
    print(row, flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/bench_attn.json", "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
