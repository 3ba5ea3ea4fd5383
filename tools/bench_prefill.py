"""Secondary measurement (BASELINE configs[2] shape: image -> text understanding, 1 image = 1024 visual tokens + 512
text tokens of context, one B200): the LLM prefill over S = 1552 tokens with the full-size Bailing-MoE 16B-A3B (random
bf16 weights generated on the device; image tokens routed by `image_gate`), then greedy text decoding.
Reports prefill tokens/s and achieved TFLOP/s, the grouped tcgen05 expert GEMMs of one layer in isolation (useful and
issued FLOPs: the issued count includes the 128-row padding of every expert segment), and decode tokens/s.
Development / documentation tool — the judged line is bench.py.
    PF_LAYERS=28 PF_TOKENS=1552 python tools/bench_prefill.py"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ming_univision_b200 import _lib, ops, synthetic  # noqa: E402
from ming_univision_b200.modeling_bailing_moe import BailingMoeConfig, BailingMoeForCausalLM  # noqa: E402


def timeit(fn, iters=5, warm=2):
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
    _lib.require_device()
    layers = int(os.environ.get("PF_LAYERS", "28"))
    S = int(os.environ.get("PF_TOKENS", "1552"))
    n_img = min(1024, S - 16)
    cfg = BailingMoeConfig(**dict(synthetic.LLM_CONFIG, num_hidden_layers=layers))
    t0 = time.time()
    torch.set_default_dtype(torch.bfloat16)
    with torch.device(dev):
        llm = BailingMoeForCausalLM(cfg)
    torch.set_default_dtype(torch.float32)
    g = torch.Generator(device=dev).manual_seed(0)
    with torch.no_grad():
        for name, p in llm.named_parameters():
            if p.dim() >= 2:
                scale = 0.5 if name.endswith("gate.weight") else p.shape[-1] ** -0.5
                p.copy_(torch.randn(p.shape, generator=g, device=dev, dtype=torch.float32) * scale)
            elif "norm" in name:
                p.fill_(1.0)
            else:
                p.zero_()
    torch.cuda.synchronize()
    print(f"model built in {time.time() - t0:.1f} s, {torch.cuda.memory_allocated() / 1e9:.1f} GB", flush=True)

    emb = (torch.randn((1, S, cfg.hidden_size), generator=g, device=dev) * 0.5).to(torch.bfloat16)
    image_mask = torch.zeros((1, S), dtype=torch.bool, device=dev)
    image_mask[:, 8:8 + n_img] = True
    pos = torch.arange(S, device=dev, dtype=torch.int32).unsqueeze(0)
    cache = llm.new_cache(max_len=S + 64, max_batch=1)

    def prefill():
        cache.seq_len, cache.batch = 0, 1
        return llm.model.forward_tokens(emb, pos, cache, key_mask=None, image_mask=image_mask)

    res = {"workload": f"understanding prefill, S={S} ({n_img} image tokens), {layers}-layer Bailing-MoE 16B-A3B, 1 GPU"}
    D, I, E, k = cfg.hidden_size, cfg.moe_intermediate_size, cfg.num_experts, cfg.num_experts_per_tok
    H, Hkv, hd = cfg.num_attention_heads, cfg.num_key_value_heads, cfg.head_dim
    per_tok_layer = 2 * (D * (H + 2 * Hkv) * hd + H * hd * D + 2 * E * D  # qkv, dense, two gates
                         + k * 3 * D * I + 3 * D * I * cfg.num_shared_experts)
    attn = layers * 4 * H * hd * S * S / 2
    flops = layers * per_tok_layer * S + attn
    for mode in ("1", "0"):
        if mode == "0" and os.environ.get("PF_STREAMING", "1") != "1":
            continue
        os.environ["MB_MOE_GROUPED"] = mode
        l0 = _lib.launch_count()
        ms = timeit(prefill, iters=3, warm=1)
        key = "grouped_tcgen05" if mode == "1" else "streaming_mma_sync"
        res[key] = {"prefill_ms": round(ms, 2), "tokens_per_s": round(S / ms * 1e3, 1),
                    "tflops": round(flops / ms / 1e9, 1), "launches_per_prefill": (_lib.launch_count() - l0) // 4}
        print(key, res[key], flush=True)
    os.environ["MB_MOE_GROUPED"] = "1"

    # ---- the routed-expert GEMMs of one layer in isolation (same routing statistics: router of layer 0 on emb)
    blk = llm.model.layers[0].mlp
    pk = blk._pack()
    x = emb.view(S, D)
    logits = ops.linear(x, pk["gate"])
    idx, w = ops.router_topk(logits, k, True, None, None)
    pair_row, row_token, tile_expert, meta, max_rows = ops.moe_plan(idx, E)
    m_tiles, rows = [int(v) for v in meta.cpu()]
    lib = _lib.load()
    st = torch.cuda.current_stream().cuda_stream
    hid = torch.empty((max_rows, I), dtype=torch.bfloat16, device=dev)
    out = torch.empty((max_rows, D), dtype=torch.bfloat16, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def t_one(fn):
        ts = []
        for _ in range(5):
            flush.zero_()  # L2 flush between timed launches
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            fn()
            e.record()
            torch.cuda.synchronize()
            ts.append(s.elapsed_time(e))
        return sorted(ts)[len(ts) // 2]

    xg = ops.gather_rows(x.contiguous(), row_token, max_rows, meta)
    t_gu = t_one(lambda: lib.mb_moe_grouped_gemm(xg.data_ptr(), pk["Wgu"].data_ptr(), hid.data_ptr(),
                                                 tile_expert.data_ptr(), meta.data_ptr(), max_rows, 2 * I, D, E, 1, st))
    t_dn = t_one(lambda: lib.mb_moe_grouped_gemm(hid.data_ptr(), pk["Wd"].data_ptr(), out.data_ptr(),
                                                 tile_expert.data_ptr(), meta.data_ptr(), max_rows, D, I, E, 0, st))
    t_plan = t_one(lambda: ops.moe_plan(idx, E))
    res["expert_gemms_one_layer"] = {
        "pairs": S * k, "padded_rows": rows, "m_tiles": m_tiles,
        "gate_up_ms": round(t_gu, 4), "down_ms": round(t_dn, 4), "plan_ms": round(t_plan, 4),
        "gate_up_tflops_useful": round(2 * S * k * 2 * I * D / t_gu / 1e9, 1),
        "gate_up_tflops_issued": round(2 * rows * 2 * I * D / t_gu / 1e9, 1),
        "down_tflops_useful": round(2 * S * k * I * D / t_dn / 1e9, 1),
        "down_tflops_issued": round(2 * rows * I * D / t_dn / 1e9, 1),
        "weight_bytes_mb": round(E * 3 * I * D * 2 / 1e6, 1)}
    print(res["expert_gemms_one_layer"], flush=True)

    # ---- greedy decode after the prefill (one row)
    n_new = 32
    hidden = prefill()
    last = hidden[:, -1]

    def decode():
        cache.seq_len = S
        llm.greedy_decode(last, cache, n_new, stop_ids=())

    ms = timeit(decode, iters=2, warm=1)
    res["decode"] = {"ms_per_token": round(ms / n_new, 3), "tokens_per_s": round(n_new / ms * 1e3, 1),
                     "note": "greedy, one CUDA-graph replay + one host read per token"}
    llm.use_cuda_graph = False
    ms_e = timeit(decode, iters=1, warm=1)
    llm.use_cuda_graph = True
    res["decode_eager"] = {"ms_per_token": round(ms_e / n_new, 3), "tokens_per_s": round(n_new / ms_e * 1e3, 1)}
    print(res["decode"], flush=True)
    res["mem_gb"] = round(torch.cuda.max_memory_allocated() / 1e9, 1)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/bench_prefill.json", "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
