"""CPU study for SURVEY.md §8f.3 (not product code, not a kernel): what would block-scaled low-precision WEIGHTS cost
the rectified-flow head in parity?  The head streams 29 GB of bf16 weights per visual token (w12 / w3 of 12 residual
blocks, 16 Euler steps) and is HBM-bound, so FP8 weights would halve its roofline time — if the sampler tolerates them.

The fp32 CPU oracle (oracle/rf_oracle.py, pinned to the live reference) samples the default-size head (1.285 B
parameters, B = 3 CFG rows, CFG 3.0 / 1.1) with the w12 / w3 matrices of every residual block rounded to
  bf16                     (what the CUDA path streams today)
  fp8 e4m3, one scale per output row          (amax / 448)
  fp8 e4m3, one scale per 128-wide K block of every row
and prints the relative L2 error of the sampled latent against the fp32-weight sample.  Activations, adaLN, biases and
the accumulation stay fp32 in all cases: this isolates the weight format.

    python tools/study_rf_weight_formats.py        (~3 min on 8 cores; output also in profiles/r02_rf_weight_formats.json)
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ming_univision_b200 import synthetic  # noqa: E402
from oracle import rf_oracle as R  # noqa: E402


def to_bf16(w):
    return w.to(torch.bfloat16).float()


def to_fp8_rows(w):
    s = w.abs().amax(dim=1, keepdim=True).clamp_min(1e-12) / 448.0
    return (w / s).to(torch.float8_e4m3fn).float() * s


def to_fp8_blocks(w, block=128):
    n, k = w.shape
    wb = w.reshape(n, k // block, block)
    s = wb.abs().amax(dim=2, keepdim=True).clamp_min(1e-12) / 448.0
    return ((wb / s).to(torch.float8_e4m3fn).float() * s).reshape(n, k)


def main():
    torch.set_num_threads(os.cpu_count())
    cfg = synthetic.RF_CONFIG
    sd = {k: v.float() for k, v in synthetic.rf_state_dict(cfg, 0).items()}
    g = torch.Generator().manual_seed(5)
    rows = []
    for B, text_cfg, image_cfg in ((3, 3.0, 1.1), (2, 3.0, 1.0)):
        z = torch.randn((B, cfg["z_channels"]), generator=g)
        noise = torch.randn((1, cfg["target_channels"]), generator=g)
        with torch.no_grad():
            ref = R.sample(sd, z, noise, int(cfg["num_sampling_steps"]), 1.0, text_cfg, image_cfg)
        for name, fn in (("bf16", to_bf16), ("fp8_e4m3_row_scale", to_fp8_rows), ("fp8_e4m3_block128_scale", to_fp8_blocks)):
            q = dict(sd)
            for k, v in sd.items():
                if ".mlp.w12.weight" in k or ".mlp.w3.weight" in k:
                    q[k] = fn(v)
            with torch.no_grad():
                out = R.sample(q, z, noise, int(cfg["num_sampling_steps"]), 1.0, text_cfg, image_cfg)
            err = float((out - ref).norm() / ref.norm())
            rows.append({"cfg_rows": B, "weights": name, "latent_rel_l2_vs_fp32_weights": round(err, 5)})
            print(rows[-1], flush=True)
    with open(os.path.join(ROOT, "profiles", "r02_rf_weight_formats.json"), "w") as f:
        json.dump({"what": __doc__.split("\n\n")[1], "rows": rows}, f, indent=1)


if __name__ == "__main__":
    main()
