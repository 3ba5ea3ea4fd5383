"""Golden fixtures pinning oracle/rf_oracle.py against the UNMODIFIED reference RectifiedFlowLoss
(mingunivision/diff_loss_rf_swiglu.py), run in the build container:   python tests/golden/make_golden_rf.py

  rf_tiny.npz  width-128 / depth-2 / 4-step head: velocity field v(x, t, z) and sample() for B = 1, 2, 3 CFG rows
  rf_full.npz  the default-size head (1.285 B parameters: width 3072, depth 12, mult 4, 16 steps): sample() for B = 2, 3
The reference draws its noise with torch.randn inside sample(); the script re-seeds torch before each call and stores
the equivalent noise tensor so the oracle (which takes the noise as an argument) sees the same values.
"""
import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")
os.environ["XFORMERS_DISABLED"] = "1"

from ming_univision_b200 import synthetic  # noqa: E402
from oracle import ref_shims  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def build(cfg, seed):
    ref_shims.install()
    import contextlib
    import io

    from diff_loss_rf_swiglu import RectifiedFlowLoss

    with contextlib.redirect_stdout(io.StringIO()):
        m = RectifiedFlowLoss(target_channels=cfg["target_channels"], z_channels=cfg["z_channels"],
                              depth=cfg["depth"], width=cfg["width"],
                              num_sampling_steps=str(cfg["num_sampling_steps"]), mlp_mult=cfg["mlp_mult"])
    m.load_state_dict(synthetic.rf_state_dict(cfg, seed), strict=True)
    return m.float().eval()


def run_case(m, cfg, B, seed, text_cfg, image_cfg, temperature):
    g = torch.Generator().manual_seed(1000 + B)
    z = torch.randn((B, cfg["z_channels"]), generator=g)
    torch.manual_seed(seed)
    noise = torch.randn(1 if text_cfg != 1.0 else B, cfg["target_channels"])
    torch.manual_seed(seed)
    with contextlib_redirect():
        with torch.no_grad():
            x = m.sample(z, temperature=temperature, text_cfg=text_cfg, image_cfg=image_cfg)
    return z, noise, x


def contextlib_redirect():
    import contextlib
    import io

    return contextlib.redirect_stdout(io.StringIO())


def main():
    torch.set_num_threads(os.cpu_count())
    for name, cfg in (("tiny", synthetic.RF_TINY_CONFIG), ("full", synthetic.RF_CONFIG)):
        m = build(cfg, 0)
        out = {}
        if name == "tiny":
            g = torch.Generator().manual_seed(5)
            xin = torch.randn((3, cfg["target_channels"]), generator=g)
            tin = torch.tensor([1.0, 0.5, 0.0625])
            cin = torch.randn((3, cfg["z_channels"]), generator=g)
            with contextlib_redirect(), torch.no_grad():
                out["net_x"], out["net_t"], out["net_c"] = xin.numpy(), tin.numpy(), cin.numpy()
                out["net_v"] = m.net(xin, tin, cin).numpy()
        for B, tc, ic in ((1, 1.0, 1.0), (2, 3.0, 1.1), (3, 3.0, 1.1)):
            if name == "full" and B == 1:
                continue
            z, noise, x = run_case(m, cfg, B, seed=11, text_cfg=tc, image_cfg=ic, temperature=0.9)
            out[f"B{B}_z"], out[f"B{B}_noise"], out[f"B{B}_x"] = z.numpy(), noise.numpy(), x.numpy()
            out[f"B{B}_cfg"] = np.array([tc, ic, 0.9])
            print(name, "B", B, "x std", float(x.std()), "rows equal", bool(B == 1 or torch.equal(x[0], x[-1])))
        np.savez_compressed(os.path.join(OUT, f"rf_{name}.npz"), seed=0, **out)


if __name__ == "__main__":
    main()
