"""GPU checks of the 3-D multimodal RoPE kernel (mb_rope3d_kv_append), run in a process of their own by
tests/test_rope3d_gpu.py: the kernel against the oracle (itself pinned to the unmodified reference functions) and the
reference's golden vectors, then the config-gated path through BailingMoeModel.forward_tokens."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

from ming_univision_b200 import _lib, ops, synthetic  # noqa: E402
from oracle import bailing_oracle as O  # noqa: E402
from parity_metrics import rel_l2  # noqa: E402

BF16 = torch.bfloat16


def tolerance(x, want, got, pos, theta, hd, sec):
    """One bf16 ulp + 4 fp32 ulps of the rotation angle times (|x1| + |x2|) — see tests/test_rope3d_cpu.py."""
    inv_freq = 1.0 / (theta ** (torch.arange(0, hd, 2).float() / hd))
    comp = torch.tensor([0 if i < sec[0] else (1 if i < sec[0] + sec[1] else 2) for i in range(hd // 2)])
    ang = (pos[..., None].float() * inv_freq)[comp, :, :, torch.arange(hd // 2)].permute(1, 2, 0)
    ang = torch.cat([ang, ang], dim=-1)[:, None]
    big = torch.maximum(torch.maximum(got.abs(), want.abs()), torch.tensor(2.0 ** -100))
    ulp = torch.exp2(torch.floor(torch.log2(big)) - 7)
    mag = x.float().abs() + O.rotate_half(x.float()).abs()
    return ulp + mag * (ang + 1.0) * 2.0 ** -21


def main():
    _lib.require_device()
    dev = torch.device("cuda:0")
    hd, theta, sec = 128, 600000.0, (16, 24, 24)
    for B, S, H, Hkv, t0, Tmax in [(2, 5, 4, 2, 0, 16), (1, 1, 16, 4, 37, 64), (3, 1, 16, 4, 11, 32),
                                   (1, 300, 16, 4, 3, 320)]:
        gen = torch.Generator().manual_seed(B * 100 + S)
        qkv = torch.randn(B * S, (H + 2 * Hkv) * hd, generator=gen).to(BF16)
        pos = torch.stack([torch.randint(0, 3000, (B, S), generator=gen), torch.randint(0, 70, (B, S), generator=gen),
                           torch.randint(0, 70, (B, S), generator=gen)])
        x = qkv.view(B, S, H + 2 * Hkv, hd)
        q, k, v = x[:, :, :H].transpose(1, 2), x[:, :, H:H + Hkv].transpose(1, 2), x[:, :, H + Hkv:].transpose(1, 2)
        cos, sin = O.mrope_tables(hd, theta, pos)
        qe, ke = O.apply_mrope(q, k, cos, sin, sec)
        want_q, want_k = qe.to(BF16).float(), ke.to(BF16).float()
        kc = torch.full((B, Hkv, Tmax, hd), 7.0, device=dev).to(BF16)
        vc = torch.full((B, Hkv, Tmax, hd), 7.0, device=dev).to(BF16)
        pos_dev = pos.reshape(3, -1).to(torch.int32).to(dev).contiguous()
        for t_dev in (None, torch.tensor([t0], dtype=torch.int32, device=dev)):
            q_out = ops.rope3d_kv_append(qkv.to(dev), pos_dev, kc, vc, B, S, H, 0 if t_dev is not None else t0, theta,
                                         sec, t_dev)
            torch.cuda.synchronize()
            got_q = q_out.float().cpu().view(B, S, H, hd).transpose(1, 2)
            got_k = kc.float().cpu()[:, :, t0:t0 + S]
            for got, want, xin, what in ((got_q, want_q, q, "q"), (got_k, want_k, k, "k")):
                tol = tolerance(xin, want, got, pos, theta, hd, sec)
                bad = (got - want).abs() > tol
                assert not bool(bad.any()), (what, B, S, float(((got - want).abs() / tol).max()))
                assert float((got != want).float().mean()) < 0.05, what
            assert torch.equal(vc.float().cpu()[:, :, t0:t0 + S], v.float()), "v is a plain copy"
            untouched = torch.ones(Tmax, dtype=torch.bool)
            untouched[t0:t0 + S] = False
            assert bool((kc.float().cpu()[:, :, untouched] == 7.0).all()), "wrote outside the appended slots"
        print("rope3d ok", (B, S, H, Hkv, t0, Tmax), flush=True)

    # golden vectors of the reference functions
    g = np.load(os.path.join(ROOT, "tests", "golden", "rope3d.npz"))
    B, S, H, Hkv, hd = (int(t) for t in g["dims"])
    q, k = torch.from_numpy(g["q"]).to(BF16), torch.from_numpy(g["k"]).to(BF16)
    qkv = torch.cat([q.transpose(1, 2), k.transpose(1, 2), torch.zeros_like(k).transpose(1, 2)], dim=2).reshape(B * S, -1)
    kc = torch.zeros((B, Hkv, S, hd), dtype=BF16, device=dev)
    vc = torch.zeros_like(kc)
    pos = torch.from_numpy(g["pos"])
    q_out = ops.rope3d_kv_append(qkv.to(dev), pos.reshape(3, -1).to(torch.int32).to(dev).contiguous(), kc, vc, B, S, H,
                                 0, float(g["theta"]))
    got_q = q_out.float().cpu().view(B, S, H, hd).transpose(1, 2)
    for got, want, xin in ((got_q, torch.from_numpy(g["q_rot"]), q), (kc.float().cpu(), torch.from_numpy(g["k_rot"]), k)):
        want16 = want.to(BF16).float()
        assert not bool(((got - want16).abs() > tolerance(xin, want16, got, pos, float(g["theta"]), hd, (16, 24, 24))).any())
    print("golden ok", flush=True)

    # config-gated model path: with all three components equal, M-RoPE rotates by the same angles as the 1-D legacy
    # path; only the rounding differs (fp32 tables and one rounding vs bf16 tables and bf16 products)
    from ming_univision_b200.modeling_bailing_moe import BailingMoeConfig, BailingMoeForCausalLM

    cfg = dict(synthetic.LLM_TINY_CONFIG)
    sd = {k: v.to(dev) for k, v in synthetic.llm_state_dict(cfg, None, None, 0).items()}
    hidden = {}
    for name, rs in (("1d", None), ("3d", {"type": "3D", "factor": 1.0})):
        with torch.device(dev):
            llm = BailingMoeForCausalLM(BailingMoeConfig(**dict(cfg, rope_scaling=rs)))
        llm.load_state_dict(sd, strict=False)
        llm = llm.to(BF16)
        ids = torch.randint(0, cfg["vocab_size"], (1, 12), generator=torch.Generator().manual_seed(3)).to(dev)
        cache = llm.new_cache(max_len=32)
        pos1 = torch.arange(12, device=dev, dtype=torch.int32).unsqueeze(0)
        pos_in = pos1 if rs is None else pos1.unsqueeze(0).expand(3, 1, 12).contiguous()
        h = llm.model.forward_tokens(llm.model.embed(ids), pos_in, cache)
        step = llm.model.embed(ids[:, :1])
        p1 = torch.full((1, 1), 12, device=dev, dtype=torch.int32)
        h2 = llm.model.forward_tokens(step, p1 if rs is None else p1.unsqueeze(0).expand(3, 1, 1).contiguous(), cache)
        hidden[name] = (h.float().cpu(), h2.float().cpu())
        if rs is None:  # 3-D ids without the config switch are refused
            try:
                llm.model.forward_tokens(step, p1.unsqueeze(0).expand(3, 1, 1).contiguous(), cache)
                raise AssertionError("3-D position ids accepted without rope_scaling.type == '3D'")
            except ValueError:
                pass
    # (bf16 tables + bf16 products on one side, fp32 tables + one rounding on the other: a bf16-sized difference)
    assert rel_l2(hidden["3d"][0], hidden["1d"][0]) < 3e-2 and rel_l2(hidden["3d"][1], hidden["1d"][1]) < 3e-2
    print("model path ok", flush=True)
    print("rope3d worker ok", flush=True)


if __name__ == "__main__":
    main()
