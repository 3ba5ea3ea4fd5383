"""Golden vectors of the 3-D multimodal RoPE (SURVEY.md row a14, config-gated variant), produced by the UNMODIFIED
reference functions — BailingMoe3DRotaryEmbedding.forward (modeling_bailing_moe.py:413-425) and
apply_multimodal_rotary_pos_emb (:463-469) — imported from /root/reference in the build container:

    python tests/golden/make_golden_rope3d.py        -> tests/golden/rope3d.npz
"""
import importlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ref_shims  # noqa: E402


def main():
    ref_shims.install()
    m = importlib.import_module("modeling_bailing_moe")
    B, S, H, Hkv, hd, theta = 2, 5, 4, 2, 128, 600000.0
    g = torch.Generator().manual_seed(31)
    q = torch.randn(B, H, S, hd, generator=g).to(torch.bfloat16)
    k = torch.randn(B, Hkv, S, hd, generator=g).to(torch.bfloat16)
    pos = torch.stack([torch.randint(0, 4000, (B, S), generator=g), torch.randint(0, 64, (B, S), generator=g),
                       torch.randint(0, 64, (B, S), generator=g)])
    rot = m.BailingMoe3DRotaryEmbedding(hd, max_position_embeddings=4096, base=theta)
    cos, sin = rot(k, position_ids=pos)
    qe, ke = m.apply_multimodal_rotary_pos_emb(q, k, cos, sin)
    assert qe.dtype == torch.float32  # the variant's contract: fp32 results, rounded once by the attention
    np.savez_compressed(os.path.join(HERE, "rope3d.npz"), q=q.float().numpy(), k=k.float().numpy(), pos=pos.numpy(),
                        q_rot=qe.numpy(), k_rot=ke.numpy(), theta=np.float32(theta), dims=np.array([B, S, H, Hkv, hd]))
    print("wrote rope3d.npz", os.path.getsize(os.path.join(HERE, "rope3d.npz")))


if __name__ == "__main__":
    main()
