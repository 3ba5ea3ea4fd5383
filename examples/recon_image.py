"""Image -> MingTok latent tokens -> image, the flow of the reference's reconstruction demo
(mingunivision/test_infer_recon_image.py) on the B200-native path: decoded u8 pixels go to the GPU, where the resize /
centre crop / normalisation (bit-exact with the PIL + torchvision stack), the encoder, the semantic decoder, the pixel
decoder and the conversion back to u8 all run; only the u8 result returns.

    python examples/recon_image.py --model <MingTok-Vision checkpoint dir> --image in.png --out recon.png [--size 512]

Without --model a seeded random-weight model of the full architecture is built (the picture is then noise — useful
only to exercise the path, e.g. on a fresh GPU box with no checkpoint)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ming_univision_b200 import synthetic  # noqa: E402
from ming_univision_b200.mingtok import MingTok, MingTokConfig  # noqa: E402
from ming_univision_b200.mingtok.utils import CenterCropProcessor, tensor_to_pil  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default=None, help="HF-layout MingTok-Vision directory (config.json + safetensors)")
    ap.add_argument("--image", required=True)
    ap.add_argument("--out", default="recon.png")
    ap.add_argument("--size", type=int, default=512, help="the demo's CenterCropProcessor(image_size=512)")
    args = ap.parse_args()
    from PIL import Image

    dev = torch.device("cuda:0")
    if args.model is not None:
        model = MingTok.from_pretrained(args.model)
    else:
        cfg = synthetic.MINGTOK_CONFIG
        with torch.device(dev):
            model = MingTok(MingTokConfig(**cfg))
        model.load_state_dict({k: v.to(dev) for k, v in synthetic.mingtok_state_dict(cfg, 0).items()}, strict=True)
    model = model.to(dev).to(torch.bfloat16)

    image = Image.open(args.image).convert("RGB")
    processor = CenterCropProcessor(image_size=args.size, mean=[0.5, 0.5, 0.5], std=[0.5, 0.5, 0.5])
    x = processor(image).unsqueeze(0)                      # [1, 3, S, S] fp32, already on the device
    out = model.forward_enc_dec(x)                         # [1, 3, S, S] in [-1, 1]
    tensor_to_pil(out).save(args.out)
    print(f"{args.image} -> {args.out} ({args.size}x{args.size}, {(args.size // 32) ** 2} latent tokens)")


if __name__ == "__main__":
    main()
