"""clock64 stamps of the first CTA of the tcgen05 GEMM (development aid): per tile, MMA role [start, accumulator free,
first operands landed, all MMAs issued] and epilogue warp [start waiting, accumulator full, epilogue done]."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ming_univision_b200 import _lib, ops
dev = torch.device("cuda:0")
for name, M, N, K, epi in [("pix.proj", 16384, 1024, 1024, ops.EPI_RESIDUAL), ("pix.fc1", 16384, 4096, 1024, ops.EPI_GELU),
                           ("pix.qkv", 16384, 3072, 1024, ops.EPI_BIAS), ("sem.proj", 4160, 1024, 1024, ops.EPI_RESIDUAL)]:
    x = torch.randn((M, K), device=dev).to(torch.bfloat16)
    w = (torch.randn((N, K), device=dev) / K ** 0.5).to(torch.bfloat16)
    b = torch.randn((N,), device=dev).to(torch.bfloat16)
    r = torch.randn((M, N), device=dev).to(torch.bfloat16)
    kw = dict(epi=epi, residual=r if epi == ops.EPI_RESIDUAL else None)
    for _ in range(3):
        ops.linear(x, w, b, **kw)
    buf = torch.zeros((16 * 8,), dtype=torch.int64, device=dev)
    _lib.load().mb_gemm_set_debug(buf.data_ptr())
    ops.linear(x, w, b, **kw)
    torch.cuda.synchronize()
    _lib.load().mb_gemm_set_debug(None)
    t = buf.cpu().view(16, 8)
    base = int(t[0, 0])
    print(name, "  tile | mma: start acc_free ops_in issued | epi: wait full done")
    for i in range(6):
        if int(t[i, 0]) == 0:
            break
        print(f"   {i:2d}   " + " ".join(f"{int(t[i, j]) - base:8d}" for j in range(7)))
