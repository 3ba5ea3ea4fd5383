"""GPU checks of the image pre- / post-processing kernels, run in a process of their own by
tests/test_preprocess_gpu.py (a fault in a first hardware run must not poison the CUDA context of the main test
process).  Every comparison is exact: u8 images and fp32 tensors bit for bit against the oracle, the committed golden
vectors and — where Pillow / torchvision are importable on the box — the libraries themselves."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from ming_univision_b200 import _lib, ops, synthetic  # noqa: E402
from ming_univision_b200.mingtok.utils import CenterCropProcessor, ResizeProcessor, tensor_to_pil  # noqa: E402
from oracle import preprocess_oracle as po  # noqa: E402

HALF = (0.5, 0.5, 0.5)
CLIP_MEAN, CLIP_STD = (0.48145466, 0.4578275, 0.40821073), (0.26862954, 0.26130258, 0.27577711)


def photo(rng, h, w):
    yy, xx = np.mgrid[0:h, 0:w]
    waves = [127 + 100 * np.sin(xx / (3.0 + c) + yy / (5.0 - c) + c) for c in range(3)]
    img = np.stack(waves, axis=2) + rng.integers(-24, 25, (h, w, 3))
    return np.clip(img, 0, 255).astype(np.uint8)


def main():
    _lib.require_device()
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(99)

    # 1. parity with the oracle over the geometry classes (both passes, one pass, none; crops of rows / columns;
    #    enlarging; a reduction large enough to narrow the horizontal tile; odd byte alignments)
    cases = [(517, 389, 256, 256, HALF, HALF), (300, 451, 512, 512, HALF, HALF), (350, 350, (512, 512), None, HALF, HALF),
             (256, 256, 256, 256, HALF, HALF), (320, 240, (224, 224), None, CLIP_MEAN, CLIP_STD),
             (97, 131, 64, 64, HALF, HALF), (131, 97, 64, 64, HALF, HALF), (600, 800, 256, 256, HALF, HALF),
             (64, 640, 64, 64, HALF, HALF), (640, 64, 64, 64, HALF, HALF), (33, 33, (7, 5), None, HALF, HALF),
             (64, 100, (64, 50), None, HALF, HALF), (100, 64, (50, 64), None, HALF, HALF),
             (40, 2600, (40, 3), None, HALF, HALF), (3, 301, (5, 299), None, HALF, HALF)]
    for h, w, size, crop, mean, std in cases:
        imgs = np.stack([photo(rng, h, w) for _ in range(3)])
        d = torch.from_numpy(imgs).to(dev)
        got = ops.image_preprocess(d, size, crop, mean, std, torch.float32)
        got16 = ops.image_preprocess(d, size, crop, mean, std, torch.bfloat16)
        torch.cuda.synchronize()
        for i in range(3):
            ref = torch.from_numpy(po.preprocess(imgs[i], size, crop, mean, std))
            assert torch.equal(got[i].cpu(), ref), ("fp32 mismatch", h, w, size, crop, i)
            assert torch.equal(got16[i].cpu(), ref.to(torch.bfloat16)), ("bf16 mismatch", h, w, size, crop, i)
        u8 = ops.image_resize_u8(d, size, crop)  # the resized + cropped image itself (out_kind 2)
        rh, rw = po.resized_output_size(h, w, size)
        want = po.resize_bicubic_u8(imgs[0], rh, rw)
        if crop is not None:
            t, l = po.center_crop_offsets(rh, rw, crop, crop)
            want = want[t:t + crop, l:l + crop]
        assert u8.dtype == torch.uint8 and np.array_equal(u8[0].cpu().numpy(), want), ("u8 mismatch", h, w, size, crop)
        print("preprocess ok", (h, w, size, crop), flush=True)

    # empty batch: no launch, empty result
    empty = ops.image_preprocess(torch.zeros((0, 50, 60, 3), dtype=torch.uint8, device=dev), 32, 32)
    assert empty.shape == (0, 3, 32, 32)
    assert ops.image_postprocess(torch.zeros((0, 3, 8, 8), device=dev)).shape == (0, 8, 8, 3)

    # 2. golden vectors produced by Pillow / torchvision
    g = np.load(os.path.join(ROOT, "tests", "golden", "preprocess.npz"))
    for c in json.loads(str(g["cases"])):
        size = tuple(c["size"]) if isinstance(c["size"], list) else c["size"]
        got = ops.image_preprocess(torch.from_numpy(g[c["name"] + ".src"]).to(dev), size, c["crop"], c["mean"], c["std"])
        assert np.array_equal(got[0].cpu().numpy(), g[c["name"] + ".tensor"]), ("golden", c["name"])
        # the u8 image behind it: invert the normalisation exactly
        u8 = ops.image_postprocess(got, c["mean"], c["std"])
        back = np.round((got[0].cpu().numpy().transpose(1, 2, 0) * np.float32(c["std"]) + np.float32(c["mean"])) * 255)
        assert np.array_equal(back.astype(np.uint8), g[c["name"] + ".u8"]), ("golden u8", c["name"])
        assert u8.shape == (1,) + g[c["name"] + ".u8"].shape
    x = torch.from_numpy(g["post.x"]).to(dev)
    assert np.array_equal(ops.image_postprocess(x)[0].cpu().numpy(), g["post.u8"]), "golden post"
    print("golden ok", flush=True)

    # 3. post-processing: every u8 level boundary, fp32 and bf16 inputs, batch
    levels = torch.arange(256, dtype=torch.float32) / 255 * 2 - 1
    grid = torch.stack([levels, torch.nextafter(levels, torch.tensor(2.0)), torch.nextafter(levels, torch.tensor(-2.0))])
    x = torch.cat([grid.clamp(-1, 1).reshape(1, 3, 16, 16), torch.rand(1, 3, 16, 16) * 2 - 1]).to(dev)
    for t in (x, x.to(torch.bfloat16)):
        got = ops.image_postprocess(t).cpu().numpy()
        for i in range(2):
            assert np.array_equal(got[i], po.postprocess(t[i].float().cpu().numpy())), ("post", t.dtype, i)
    print("postprocess ok", flush=True)

    # 4. the drop-in classes against the torchvision stack they replace (reference composition)
    try:
        import torchvision.transforms as T
        from PIL import Image
        from torchvision.transforms import InterpolationMode
    except ImportError:
        print("torchvision / Pillow not importable here: class checks against the oracle only", flush=True)
        T = None
    img = photo(rng, 413, 620)
    for proc, size, crop in ((CenterCropProcessor(image_size=256, mean=[0.5] * 3, std=[0.5] * 3), 256, 256),
                             (CenterCropProcessor.from_config({"image_size": 128}), 128, 128),
                             (ResizeProcessor(image_size=224), (224, 224), None)):
        item = Image.fromarray(img) if T is not None else img
        got = proc(item)
        assert got.is_cuda and got.dtype == torch.float32 and got.dim() == 3
        ref = po.preprocess(img, size, crop, proc.mean, proc.std)
        assert np.array_equal(got.cpu().numpy(), ref), ("processor vs oracle", size, crop)
        if T is not None:
            tf = [T.Resize(size=size, interpolation=InterpolationMode.BICUBIC)]
            if crop is not None:
                tf.append(T.CenterCrop(crop))
            tv = T.Compose(tf + [T.ToTensor(), T.Normalize(proc.mean, proc.std)])(item)
            assert torch.equal(got.cpu(), tv), ("processor vs torchvision", size, crop)
    if T is not None:
        y = torch.rand(1, 3, 40, 56, device=dev) * 2 - 1
        half = torch.tensor(HALF, device=dev).view(1, -1, 1, 1)
        assert np.array_equal(np.asarray(tensor_to_pil(y)), np.asarray(T.ToPILImage()((y * half + half)[0])))
        # a photograph-sized input against Pillow itself (the oracle is not involved)
        big = photo(rng, 1536, 2048)
        got = ops.image_preprocess(torch.from_numpy(big).to(dev), (1024, 1024), None)
        u8 = ops.image_postprocess(got)[0].cpu().numpy()
        tvbig = T.Compose([T.Resize((1024, 1024), interpolation=InterpolationMode.BICUBIC), T.ToTensor(),
                           T.Normalize(HALF, HALF)])(Image.fromarray(big))
        assert torch.equal(got[0].cpu(), tvbig), "1536x2048 -> 1024x1024 vs torchvision"
        assert u8.shape == (1024, 1024, 3)
    print("processors ok", flush=True)

    # 5. configs[1]-shaped batch (64 images -> 256 x 256) and the reconstruction demo's flow
    #    (mingunivision/test_infer_recon_image.py:15-28): processor -> forward_enc_dec -> tensor_to_pil
    batch = torch.from_numpy(np.stack([photo(rng, 300, 400) for _ in range(4)])).to(dev).repeat(16, 1, 1, 1)
    out = CenterCropProcessor(image_size=256).batch(batch)
    assert out.shape == (64, 3, 256, 256) and torch.equal(out[:4], out[60:])
    assert np.array_equal(out[1].cpu().numpy(), po.preprocess(batch[1].cpu().numpy(), 256, 256))
    from ming_univision_b200.mingtok import MingTok, MingTokConfig

    cfg = synthetic.MINGTOK_TINY_CONFIG
    with torch.device(dev):
        model = MingTok(MingTokConfig(**cfg))
    model.load_state_dict({k: v.to(dev) for k, v in synthetic.mingtok_state_dict(cfg, 0).items()}, strict=True)
    model = model.to(torch.bfloat16)
    image = CenterCropProcessor(image_size=128)(photo(rng, 200, 150)).cuda().unsqueeze(0)
    recon = model.forward_enc_dec(image)
    pil = tensor_to_pil(recon)
    assert pil.size == (128, 128) and pil.mode == "RGB"
    # the fused tail: head rows -> u8 pixels in one pass == unpatchify_clamp followed by the u8 conversion
    feats = model.forward(image)["x_norm_patchtokens"]
    fused = model.forward_pixel_decoder(feats, out_dtype=torch.uint8)
    for dt in (torch.bfloat16, torch.float32):
        assert torch.equal(fused, ops.image_postprocess(model.forward_pixel_decoder(feats, out_dtype=dt))), dt
    assert fused.shape == (1, 128, 128, 3) and np.array_equal(fused[0].cpu().numpy(), np.asarray(pil))
    torch.cuda.synchronize()
    print("demo flow ok", flush=True)
    print("preprocess worker ok", flush=True)


if __name__ == "__main__":
    main()
