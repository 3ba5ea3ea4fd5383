"""Generates the golden fixtures that pin oracle/ against the UNMODIFIED reference (run in the BUILD container,
where /root/reference exists; the GPU box only reads the committed .npz files).

    python tests/golden/make_golden.py

For each case the reference's own classes (mingtok/modeling_mingtok.py:MingTok, built through oracle/ref_shims.py,
fp32, eager attention) are loaded with ming_univision_b200.synthetic weights and run on synthetic images; inputs are
regenerated from seeds by the tests, outputs are stored.
  mingtok_tiny_*.npz   full tensors of a 2+2+2-layer, 128-wide MingTok at 128x128 and 64x64 (pos-embed interpolation)
  mingtok_full_256.npz strided samples + statistics of the FULL-SIZE model (697.7 M params) at 1x3x256x256
"""
import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")

from ming_univision_b200 import synthetic  # noqa: E402
from oracle import ref_shims  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def run_reference(config, seed, img, steps=0):
    sd = synthetic.mingtok_state_dict(config, seed)
    model = ref_shims.build_reference_mingtok(config, sd, fa_enable=False)
    with torch.no_grad():
        out = model.forward(img)
        recon = model.forward_pixel_decoder(out["x_norm_patchtokens"])
        res = {"latent": out["latent"], "feats": out["x_norm_patchtokens"], "recon": recon}
        if steps:
            # incremental semantic decoding through the reference's own KV cache (modeling_mingtok.py:165-174)
            lat = out["latent"][:, :steps]
            pkv, outs = None, []
            for t in range(steps):
                r = model.forward_feature_decoder(lat[:, t:t + 1], past_key_values=pkv)
                pkv = r["past_key_values"]
                outs.append(r["x_norm_patchtokens"])
            res["feats_incremental"] = torch.cat(outs, dim=1)
    return res


def main():
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    tiny = synthetic.MINGTOK_TINY_CONFIG
    for size, batch in ((128, 2), (64, 3)):
        img = synthetic.synthetic_images(batch, size, seed=1234)
        res = run_reference(tiny, 0, img, steps=(size // 32) ** 2)
        np.savez_compressed(os.path.join(OUT, f"mingtok_tiny_{size}.npz"), seed=0, img_seed=1234, batch=batch,
                            size=size, **{k: v.numpy() for k, v in res.items()})
        print("tiny", size, {k: tuple(v.shape) for k, v in res.items()})

    full = synthetic.MINGTOK_CONFIG
    img = synthetic.synthetic_images(1, 256, seed=1234)
    res = run_reference(full, 0, img, steps=4)
    packed = {}
    for k, v in res.items():
        flat = v.flatten()
        idx = torch.arange(0, flat.numel(), max(1, flat.numel() // 4096))[:4096]
        packed[k + "_idx"] = idx.numpy()
        packed[k + "_val"] = flat[idx].numpy()
        packed[k + "_stats"] = np.array([flat.mean().item(), flat.std().item(), flat.abs().max().item()])
    np.savez_compressed(os.path.join(OUT, "mingtok_full_256.npz"), seed=0, img_seed=1234, **packed)
    print("full", {k: tuple(v.shape) for k, v in res.items()})


if __name__ == "__main__":
    main()
