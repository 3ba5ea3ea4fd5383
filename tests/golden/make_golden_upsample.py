"""Golden fixtures for the UPSAMPLED position table (inputs larger than the model's native resolution — the understanding
path, BASELINE configs[2]: 1024 x 1024 -> 32 x 32 patches + cls = 1025 encoder / semantic-decoder tokens on a table
trained for 16 x 16; mingtok/vision_transformer/vision_transformer.py:183-215), from the UNMODIFIED reference run in the
build container:

    python tests/golden/make_golden_upsample.py   ->   tests/golden/mingtok_tiny_256.npz, mingtok_full_1024.npz

  mingtok_tiny_256.npz    full tensors of the tiny MingTok (native 128 px: 4 x 4 table) at 1 x 3 x 256 x 256 (8 x 8)
  mingtok_full_1024.npz   strided samples + statistics of the FULL-SIZE encoder + semantic decoder at 1 x 3 x 1024 x 1024
"""
import os
import sys
import time
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")

from ming_univision_b200 import synthetic  # noqa: E402
from oracle import ref_shims  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    torch.set_num_threads(os.cpu_count())
    tiny = synthetic.MINGTOK_TINY_CONFIG
    model = ref_shims.build_reference_mingtok(tiny, synthetic.mingtok_state_dict(tiny, 0), fa_enable=False)
    img = synthetic.synthetic_images(1, 256, seed=1234)
    with torch.no_grad():
        out = model.forward(img)
        recon = model.forward_pixel_decoder(out["x_norm_patchtokens"])
    np.savez_compressed(os.path.join(OUT, "mingtok_tiny_256.npz"), seed=0, img_seed=1234, batch=1, size=256,
                        latent=out["latent"].numpy(), feats=out["x_norm_patchtokens"].numpy(), recon=recon.numpy())
    print("tiny 256", tuple(out["latent"].shape), tuple(out["x_norm_patchtokens"].shape), tuple(recon.shape))

    full = synthetic.MINGTOK_CONFIG
    t0 = time.time()
    model = ref_shims.build_reference_mingtok(full, synthetic.mingtok_state_dict(full, 0), fa_enable=False)
    img = synthetic.synthetic_images(1, 1024, seed=1234)
    with torch.no_grad():
        out = model.forward(img)
    packed = {}
    for k, v in (("latent", out["latent"]), ("feats", out["x_norm_patchtokens"])):
        flat = v.flatten()
        idx = torch.arange(0, flat.numel(), max(1, flat.numel() // 8192))[:8192]
        packed[k + "_idx"] = idx.numpy()
        packed[k + "_val"] = flat[idx].numpy()
        packed[k + "_shape"] = np.array(v.shape)
        packed[k + "_stats"] = np.array([flat.mean().item(), flat.std().item(), flat.abs().max().item()])
    np.savez_compressed(os.path.join(OUT, "mingtok_full_1024.npz"), seed=0, img_seed=1234, **packed)
    print("full 1024", tuple(out["latent"].shape), tuple(out["x_norm_patchtokens"].shape), f"{time.time() - t0:.0f} s")


if __name__ == "__main__":
    main()
