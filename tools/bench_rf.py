"""Micro-benchmark of the rectified-flow head: weight-streaming kernel GB/s and RectifiedFlowLoss.sample latency
(default-size head: width 3072, depth 12, mult 4, 16 steps).  Development tool."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ming_univision_b200 import ops, synthetic  # noqa: E402
from ming_univision_b200.diff_loss_rf_swiglu import RectifiedFlowLoss  # noqa: E402

dev = torch.device("cuda:0")
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def timeit(fn, iters=10, warm=3, do_flush=True):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        if do_flush:
            flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    res = []
    for name, M, N, K, epi in [("w12", 3, 16384, 3072, ops.EPI_SWIGLU), ("w3", 3, 3072, 8192, ops.EPI_BIAS),
                               ("w12.b2", 2, 16384, 3072, ops.EPI_SWIGLU), ("ada", 3, 9216, 3072, ops.EPI_BIAS),
                               ("lm_head", 3, 126464, 2048, ops.EPI_BIAS), ("qkv", 3, 3072, 2048, ops.EPI_BIAS)]:
        x = torch.randn((M, K), device=dev).to(torch.bfloat16)
        w = (torch.randn((N, K), device=dev) / K ** 0.5).to(torch.bfloat16)
        b = torch.randn((N,), device=dev).to(torch.bfloat16)
        ms = timeit(lambda: ops.gemv(x, w, b, epi=epi))
        res.append({"op": "gemv." + name, "M": M, "N": N, "K": K, "ms": round(ms, 4),
                    "gbs": round(N * K * 2 / ms / 1e6, 1)})
        print(res[-1], flush=True)
    cfg = synthetic.RF_CONFIG
    with torch.device(dev):
        m = RectifiedFlowLoss(cfg["target_channels"], cfg["z_channels"], cfg["depth"], cfg["width"],
                              str(cfg["num_sampling_steps"]), mlp_mult=cfg["mlp_mult"])
    m.load_state_dict({k: v.to(dev) for k, v in synthetic.rf_state_dict(cfg, 0).items()})
    m = m.to(torch.bfloat16)
    weight_bytes = sum(p.numel() for p in m.parameters()) * 2
    ada = sum(p.numel() for n, p in m.named_parameters() if "adaLN" in n) * 2
    for fused in ("1", "0"):
        os.environ["MB_RF_FUSED"] = fused
        m._graphs = {}
        for B in (2, 3):
            z = torch.randn((B, cfg["z_channels"]), device=dev)
            ms = timeit(lambda: m.sample(z, temperature=1.0, text_cfg=3.0, image_cfg=1.1), iters=5, do_flush=False)
            alg = 16 * (weight_bytes - ada) + ada  # bytes that must stream per token after the adaLN hoist
            res.append({"op": f"rf.sample.B{B}.{'persistent' if fused == '1' else 'layers'}", "ms": round(ms, 3),
                        "tokens_per_s": round(1e3 / ms, 1), "alg_gb": round(alg / 1e9, 2),
                        "gbs": round(alg / ms / 1e6, 1), "naive_alg_gb": round(16 * weight_bytes / 1e9, 2)})
            print(res[-1], flush=True)
    os.environ["MB_RF_FUSED"] = "1"
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/bench_rf.json", "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
