"""Micro-benchmarks of the C-ABI operators on the MingTok config-2 shapes (B=64, 256x256): CUDA-event timing,
L2 flushed between iterations; prints achieved TFLOP/s (GEMM, attention) or GB/s (row kernels) next to cuBLAS /
torch for orientation.  Development tool, not the judged benchmark (that is bench.py)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ming_univision_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
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
    shapes = [  # (name, M, N, K, epi)
        ("enc.qkv", 4160, 2304, 768, "bias"), ("enc.proj", 4160, 768, 768, "residual"),
        ("enc.w12", 4160, 4096, 768, "swiglu"), ("enc.w3", 4160, 768, 2048, "residual"),
        ("sem.qkv", 4160, 3072, 1024, "bias"), ("sem.proj", 4160, 1024, 1024, "residual"),
        ("sem.w12", 4160, 5632, 1024, "swiglu"), ("sem.w3", 4160, 1024, 2816, "residual"),
        ("pix.qkv", 16384, 3072, 1024, "bias"), ("pix.proj", 16384, 1024, 1024, "residual"),
        ("pix.fc1", 16384, 4096, 1024, "gelu"), ("pix.fc2", 16384, 1024, 4096, "residual"),
        ("patch", 4096, 768, 3072, "bias"), ("s2p", 4096, 4096, 1024, "bias"), ("sq8k", 8192, 8192, 8192, "bias"),
    ]
    for name, M, N, K, epi in shapes:
        x = torch.randn((M, K), device=dev).to(torch.bfloat16)
        w = (torch.randn((N, K), device=dev) / K ** 0.5).to(torch.bfloat16)
        b = torch.randn((N,), device=dev).to(torch.bfloat16)
        n_out = N // 2 if epi == "swiglu" else N
        r = torch.randn((M, n_out), device=dev).to(torch.bfloat16)
        out = torch.empty((M, n_out), dtype=torch.bfloat16, device=dev)
        e = {"bias": ops.EPI_BIAS, "gelu": ops.EPI_GELU, "swiglu": ops.EPI_SWIGLU, "residual": ops.EPI_RESIDUAL}[epi]
        ms_t = timeit(lambda: torch.nn.functional.linear(x, w, b))
        row = {"op": name, "M": M, "N": N, "K": K, "epi": epi,
               "cublas_tflops": round(2.0 * M * N * K / ms_t / 1e9, 1)}
        from ming_univision_b200 import _lib
        for tname, cg, bn in (("auto", 0, 0), ("pair256", 2, 256), ("pair128", 2, 128), ("single256", 1, 256),
                              ("single128", 1, 128)):
            if epi == "swiglu" and bn == 128:
                continue
            _lib.load().mb_gemm_force_tile(cg, bn)
            ms = timeit(lambda: ops.linear(x, w, b, epi=e, residual=r if epi == "residual" else None, out=out))
            row[tname] = round(2.0 * M * N * K / ms / 1e9, 1)
        _lib.load().mb_gemm_force_tile(0, 0)
        res.append(row)
        print(res[-1], flush=True)
    for name, B, S, H, causal in [("enc.attn", 64, 65, 12, False), ("sem.attn", 64, 65, 16, True),
                                  ("pix.attn", 64, 256, 16, False), ("pix512.attn", 16, 1024, 16, False)]:
        qkv = torch.randn((B, S, 3 * H * 64), device=dev).to(torch.bfloat16)
        ms = timeit(lambda: ops.attention_hd64(qkv, B, S, H, causal))
        fl = 4.0 * B * H * S * S * 64 * (0.5 if causal else 1.0)
        q, k, v = qkv.view(B, S, 3, H, 64).permute(2, 0, 3, 1, 4)
        ms_t = timeit(lambda: torch.nn.functional.scaled_dot_product_attention(q, k, v, is_causal=causal))
        res.append({"op": name, "ms": round(ms, 4), "tflops": round(fl / ms / 1e9, 1), "sdpa_ms": round(ms_t, 4)})
        print(res[-1], flush=True)
    for name, rows, dim in [("enc.ln", 4160, 768), ("pix.ln", 16384, 1024)]:
        x = torch.randn((rows, dim), device=dev).to(torch.bfloat16)
        g = torch.ones((dim,), device=dev, dtype=torch.bfloat16)
        ms = timeit(lambda: ops.layernorm(x, g, g, 1e-6, 0))
        res.append({"op": name, "ms": round(ms, 4), "gbs": round(rows * dim * 4 / ms / 1e6, 1)})
        print(res[-1], flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/bench_ops.json", "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
