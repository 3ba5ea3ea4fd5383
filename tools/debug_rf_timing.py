"""Debug aid: where the persistent rectified-flow sampler kernel spends its time (mb_rf_set_debug: wall time per phase
kind, accumulated by thread 0 of the first and the last CTA over one sample)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ming_univision_b200 import _lib, synthetic  # noqa: E402
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
m.sample(z, temperature=1.0, text_cfg=3.0, image_cfg=1.1)
buf = torch.zeros((16,), dtype=torch.int64, device=dev)
lib = _lib.load()
lib.mb_rf_set_debug(buf.data_ptr())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
m.sample(z, temperature=1.0, text_cfg=3.0, image_cfg=1.1)
e1.record()
torch.cuda.synchronize()
lib.mb_rf_set_debug(None)
names = ["adaLN prologue", "w12 stream+epi", "barrier after w12", "w3 stream+epi", "barrier after w3", "in/final/euler",
         "their barriers", "-"]
print(f"sample {e0.elapsed_time(e1):.3f} ms")
for cta, off in (("first CTA", 0), ("last CTA", 8)):
    vals = buf[off:off + 8].tolist()
    print(cta, " ".join(f"{n}: {v / 1e3:.0f} us" for n, v in zip(names, vals) if n != "-"), f"| total {sum(vals) / 1e3:.0f} us")
