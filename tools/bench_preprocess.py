"""Image pre- / post-processing: the device kernels against the CPU transforms they replace.

For every workload: device-resident time (CUDA events on the launching stream, inputs in HBM, L2 flushed between
iterations), end-to-end time from pinned host u8 pixels (H2D inside the timed region), algorithmic bytes / time against
the HBM peak, and — beside it — torchvision's own pipeline on PIL images on the host cores (the reference path,
mingtok/utils/processor.py:17-27), on a bounded sample.  Parity (exact) is asserted on the first image of every workload
before anything is timed.  One JSON line per workload.  `--cpu-only` prints the host baseline alone (no GPU needed).

    python tools/bench_preprocess.py [--cpu-only] [--iters 20]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

HALF = (0.5, 0.5, 0.5)
# (name, images per call, H, W, size, crop): the reference's three processor configurations on photograph-sized inputs
WORKLOADS = [
    ("recon_256_b64", 64, 384, 512, 256, 256),             # BASELINE configs[1]: 64 images -> 256 x 256
    ("gen_512", 1, 1536, 2048, 512, 512),                  # gen_processor: CenterCrop 512 (processing_bailingmm.py:176)
    ("und_1024", 1, 1536, 2048, (1024, 1024), None),       # vis_processor: Resize((1024, 1024)) (:175)
    ("gen_512_b16", 16, 768, 1024, 512, 512),
]


def photo(rng, h, w):
    yy, xx = np.mgrid[0:h, 0:w]
    img = np.stack([127 + 100 * np.sin(xx / (3.0 + c) + yy / (5.0 - c) + c) for c in range(3)], axis=2)
    return np.clip(img + rng.integers(-24, 25, (h, w, 3)), 0, 255).astype(np.uint8)


def cpu_pipeline(size, crop):
    import torchvision.transforms as T
    from torchvision.transforms import InterpolationMode

    tf = [T.Resize(size=size, interpolation=InterpolationMode.BICUBIC)]
    if crop is not None:
        tf.append(T.CenterCrop(crop))
    return T.Compose(tf + [T.ToTensor(), T.Normalize(HALF, HALF)])


def time_cpu(imgs, size, crop, budget_s=5.0):
    """torchvision on PIL images, one image at a time on one core (the reference's loop); bounded sample."""
    from PIL import Image

    tf = cpu_pipeline(size, crop)
    pil = [Image.fromarray(i) for i in imgs]
    tf(pil[0])
    n, t0 = 0, time.perf_counter()
    while True:
        tf(pil[n % len(pil)])
        n += 1
        if time.perf_counter() - t0 > budget_s or n >= 4 * len(pil):
            break
    return (time.perf_counter() - t0) / n, n


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cpu-only", action="store_true")
    ap.add_argument("--iters", type=int, default=20)
    args = ap.parse_args()
    rng = np.random.default_rng(0)
    if not args.cpu_only:
        from ming_univision_b200 import _lib, ops

        _lib.require_device()
        dev = torch.device("cuda:0")
        flush = torch.empty((256 << 20,), dtype=torch.uint8, device=dev)  # > 126 MB L2
        peaks_path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
        hbm_peak = None
        if os.path.exists(peaks_path):
            with open(peaks_path) as f:
                pk = json.load(f)
            hbm_peak = pk.get("hbm_gbs")  # measured copy bandwidth (driver-written)
        if hbm_peak is None:
            hbm_peak = 6489.0  # the profiling recipe's fallback, as in bench.py
    for name, n, h, w, size, crop in WORKLOADS:
        distinct = [photo(rng, h, w) for _ in range(min(n, 4))]
        imgs = np.stack([distinct[i % len(distinct)] for i in range(n)])
        cpu_s, cpu_n = time_cpu(distinct, size, crop)
        row = {"workload": name, "images": n, "in": [h, w], "size": size, "crop": crop,
               "cpu_ms_per_image": round(cpu_s * 1e3, 3), "cpu_sample_images": cpu_n, "cpu_cores": 1}
        if not args.cpu_only:
            host = torch.from_numpy(imgs).pin_memory()
            d = host.to(dev)
            out = ops.image_preprocess(d, size, crop, HALF, HALF, torch.bfloat16)
            ref = cpu_pipeline(size, crop)(__import__("PIL.Image", fromlist=["fromarray"]).fromarray(imgs[0]))
            assert torch.equal(ops.image_preprocess(d[:1], size, crop)[0].cpu(), ref), "parity"
            assert torch.equal(out[0].cpu(), ref.to(torch.bfloat16)), "parity (bf16)"
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            dev_ms = e2e_ms = 0.0
            for it in range(args.iters + 3):
                flush.fill_(it & 0xFF)
                s.record()
                ops.image_preprocess(d, size, crop, HALF, HALF, torch.bfloat16)
                e.record()
                torch.cuda.synchronize()
                if it >= 3:
                    dev_ms += s.elapsed_time(e) / args.iters
            for it in range(args.iters + 3):
                flush.fill_(it & 0xFF)
                torch.cuda.synchronize()
                s.record()
                ops.image_preprocess(host.to(dev, non_blocking=True), size, crop, HALF, HALF, torch.bfloat16)
                e.record()
                torch.cuda.synchronize()
                if it >= 3:
                    e2e_ms += s.elapsed_time(e) / args.iters
            u8 = ops.image_postprocess(out)
            post_ms = 0.0
            for it in range(args.iters + 3):
                flush.fill_(it & 0xFF)
                s.record()
                ops.image_postprocess(out)
                e.record()
                torch.cuda.synchronize()
                if it >= 3:
                    post_ms += s.elapsed_time(e) / args.iters
            rh, rw = ops.resized_output_size(h, w, size)
            # algorithmic bytes: every source pixel whose column AND row survive the crop, once, + the output once
            oh, ow = out.shape[2], out.shape[3]
            kept = (w * ow / rw) * (h * oh / rh)  # input pixels under the kept window (columns x rows)
            alg_bytes = n * (kept * 3 + oh * ow * 3 * 2)
            row.update({"gpu_ms": round(dev_ms, 4), "gpu_e2e_ms": round(e2e_ms, 4), "h2d_bytes": int(host.numel()),
                        "gpu_images_per_s": round(n / dev_ms * 1e3, 1), "e2e_images_per_s": round(n / e2e_ms * 1e3, 1),
                        "cpu_images_per_s_1core": round(1.0 / cpu_s, 1),
                        "roofline": {"bound": "hbm", "achieved": round(alg_bytes / dev_ms / 1e6, 1), "peak": hbm_peak,
                                     "unit": "GB/s", "frac": (round(alg_bytes / dev_ms / 1e6 / hbm_peak, 4)
                                                              if hbm_peak else None), "traffic": None},
                        "postprocess_ms": round(post_ms, 4), "postprocess_out_bytes": int(u8.numel())})
        print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
