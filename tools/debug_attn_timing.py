"""Phase timestamps of CTA 0 of the tcgen05 attention kernel (development aid)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ming_univision_b200 import _lib, ops
dev = torch.device("cuda:0")
B, S, H = 16, 1024, 16
qkv = torch.randn((B, S, 3 * H * 64), device=dev).to(torch.bfloat16)
ops.set_attn_backend(1)
for _ in range(3):
    ops.attention_hd64(qkv, B, S, H, False)
buf = torch.zeros((64 * 16,), dtype=torch.int64, device=dev)
_lib.load().mb_attn_set_debug(buf.data_ptr())
ops.attention_hd64(qkv, B, S, H, False)
torch.cuda.synchronize()
_lib.load().mb_attn_set_debug(None)
t = buf.cpu().view(64, 16)
base = int(t[0, 0])
names = ["sm:wait_s", "sm:got_s", "sm:-", "sm:exp0", "sm:exp1", "sm:xchg", "sm:arrive", "-", "mma:k_wait", "mma:S_iss", "mma:p_wait", "mma:p_ok", "mma:v_ok", "mma:PV_iss"]
print(" g " + " ".join(f"{n:>10s}" for n in names))
for g in range(24):
    print(f"{g:2d} " + " ".join(f"{int(t[g, i]) - base:10d}" for i in range(14)))
