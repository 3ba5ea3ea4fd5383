"""ncu helper: one eager RectifiedFlowLoss.sample (default-size head, B = 2) bracketed by cudaProfilerStart/Stop:
    ncu --profile-from-start off --metrics gpu__time_duration.sum --csv python tools/profile_rf.py"""
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
m.use_cuda_graph = False
B = int(os.environ.get("RF_ROWS", "3"))
z = torch.randn((B, cfg["z_channels"]), device=dev)
m.sample(z, temperature=1.0, text_cfg=3.0, image_cfg=1.1, groups=max(1, B // 3))
torch.cuda.synchronize()
torch.cuda.profiler.start()
m.sample(z, temperature=1.0, text_cfg=3.0, image_cfg=1.1, groups=max(1, B // 3))
torch.cuda.synchronize()
torch.cuda.profiler.stop()
