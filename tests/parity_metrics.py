"""Metrics used by the parity tests and bench.py: relative L2, PSNR, and a Fréchet distance over a fixed seeded
random-feature extractor (the offline stand-in for rFID — no Inception weights exist in this environment, SURVEY.md
§7 'hard parts')."""
import math

import numpy as np
import torch


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def psnr(a: torch.Tensor, b: torch.Tensor, data_range: float = 2.0) -> float:
    """Images in [-1, 1] (data range 2)."""
    mse = float(((a.double().cpu() - b.double().cpu()) ** 2).mean())
    return 10.0 * math.log10(data_range ** 2 / max(mse, 1e-30))


def _random_features(img: torch.Tensor, dim: int = 16, seed: int = 99) -> np.ndarray:
    """img [B,3,H,W] in [-1,1] -> [B*16, dim]: 3 fixed random conv layers, 4x4 grid of pooled descriptors per image."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    x = img.double().cpu()
    chans = [3, 16, 32, dim]
    for i in range(3):
        w = torch.randn((chans[i + 1], chans[i], 3, 3), generator=g, dtype=torch.float64) / math.sqrt(chans[i] * 9)
        x = torch.tanh(torch.nn.functional.conv2d(x, w, stride=2, padding=1))
    x = torch.nn.functional.adaptive_avg_pool2d(x, 4)  # [B, dim, 4, 4]
    return x.permute(0, 2, 3, 1).reshape(-1, dim).numpy()


def frechet_distance(img_a: torch.Tensor, img_b: torch.Tensor) -> float:
    """||mu_a - mu_b||^2 + Tr(Ca + Cb - 2 (Ca Cb)^{1/2}) over the random-feature descriptors of the two image sets."""
    from scipy import linalg

    fa, fb = _random_features(img_a), _random_features(img_b)
    mu_a, mu_b = fa.mean(0), fb.mean(0)
    ca, cb = np.cov(fa, rowvar=False), np.cov(fb, rowvar=False)
    covmean = np.asarray(linalg.sqrtm(ca.dot(cb))).real
    return float(((mu_a - mu_b) ** 2).sum() + np.trace(ca) + np.trace(cb) - 2 * np.trace(covmean))
