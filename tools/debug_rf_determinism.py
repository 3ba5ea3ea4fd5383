"""Debug aid: is the persistent RF sampler bit-reproducible (eager launches, graph replays, interleaved inputs)?"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ming_univision_b200 import synthetic  # noqa: E402
from ming_univision_b200.diff_loss_rf_swiglu import RectifiedFlowLoss  # noqa: E402

dev = torch.device("cuda:0")
cfg = synthetic.RF_CONFIG
with torch.device(dev):
    m = RectifiedFlowLoss(cfg["target_channels"], cfg["z_channels"], cfg["depth"], cfg["width"],
                          str(cfg["num_sampling_steps"]), mlp_mult=cfg["mlp_mult"])
m.load_state_dict({k: v.to(dev) for k, v in synthetic.rf_state_dict(cfg, 0).items()})
m = m.to(torch.bfloat16)
g = torch.Generator().manual_seed(1)
zs = [torch.randn((3, cfg["z_channels"]), generator=g).to(dev) for _ in range(2)]
noise = torch.randn((1, 32), generator=g).to(dev)
for direct in ("-",):
    m._graphs = {}
    outs = {}
    for mode in ("eager", "graph"):
        m.use_cuda_graph = mode == "graph"
        for rep in range(3):
            for zi, z in enumerate(zs):
                x = m.sample(z, temperature=1.0, text_cfg=3.0, image_cfg=1.1, noise=noise)
                outs.setdefault(zi, []).append((mode, rep, x.clone()))
    for zi, lst in outs.items():
        ref = lst[0][2]
        bad = [(mo, r, float((x - ref).abs().max())) for mo, r, x in lst if not torch.equal(x, ref)]
        print(f"input {zi}: {len(lst)} runs, mismatching: {bad}")
