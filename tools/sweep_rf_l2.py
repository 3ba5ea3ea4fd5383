"""Timing harness of the persistent RF sampler (csrc/rf_fused.cu) at 6, 3 and 2 rows: per setting the sample latency (CUDA
events, median of 7, eager launches incl. the adaLN hoist), the equality of the result with the first setting, and the
per-phase wall times of the first / last CTA (mb_rf_set_debug).  Used for the round-2 experiments summarised in
profiles/r02_rf_experiments.md:
  * RF_LIB=<path>: a library variant built with other compile-time constants (e.g. -DMB_RF_KC=512);
  * SWEEP="k=v,k=v;k=v;...": MB_RF_<k> environment knobs per setting — they only exist in the experimental build
    (profiles/r02_rf_experiments/producer_knobs.patch); the shipped kernel ignores them, so SWEEP=";" = two baseline runs;
  * ROWS="6x3,3x3,2x2": rows x CFG rows per launch.
Development tool; output: gpurun_out/sweep_rf_l2.json (or SWEEP_OUT)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ming_univision_b200 import _lib, synthetic  # noqa: E402
from ming_univision_b200.diff_loss_rf_swiglu import RectifiedFlowLoss  # noqa: E402

if os.environ.get("RF_LIB"):  # a library variant built with other compile-time constants (e.g. -DMB_RF_KC=512)
    _lib.LIB_PATH = os.path.abspath(os.environ["RF_LIB"])
dev = torch.device("cuda:0")
cfg = synthetic.RF_CONFIG
with torch.device(dev):
    m = RectifiedFlowLoss(cfg["target_channels"], cfg["z_channels"], cfg["depth"], cfg["width"],
                          str(cfg["num_sampling_steps"]), mlp_mult=cfg["mlp_mult"])
m.load_state_dict({k: v.to(dev) for k, v in synthetic.rf_state_dict(cfg, 0).items()})
m = m.to(torch.bfloat16)
m.use_cuda_graph = False
lib = _lib.load()
weight_bytes = sum(p.numel() for p in m.parameters()) * 2
ada = sum(p.numel() for n, p in m.named_parameters() if "adaLN" in n) * 2
alg = 16 * (weight_bytes - ada)
names = ["adaLN prologue", "w12 stream+epi", "barrier after w12", "w3 stream+epi", "barrier after w3", "in/final/euler",
         "their barriers"]
res = []
KNOBS = ("L2_AHEAD", "L2_MODE", "EVICT_FIRST", "BAR_RED", "BAR_SLEEP")  # (of the experimental build, see profiles/)
default = ["", "BAR_RED=1", "BAR_RED=1,BAR_SLEEP=20", "BAR_RED=1,BAR_SLEEP=60", "BAR_SLEEP=40", "EVICT_FIRST=1"]
for mode in (0, 1):
    for ev in (0, 1):
        for ahead in (1, 2, 4, 8):
            default.append(f"L2_AHEAD={ahead},L2_MODE={mode},EVICT_FIRST={ev}")
default += ["BAR_RED=1,EVICT_FIRST=1,L2_AHEAD=2,L2_MODE=1", ""]
sweep = os.environ["SWEEP"].split(";") if os.environ.get("SWEEP") is not None else default
row_sets = [tuple(int(v) for v in r.split("x")) for r in os.environ.get("ROWS", "6x3,3x3,2x2").split(",")]
for B, cfg_rows in row_sets:
    z = torch.randn((B, cfg["z_channels"]), device=dev)
    ref = None
    for setting in sweep:
        kv = dict(item.split("=") for item in setting.split(",") if item)
        for k in KNOBS:
            os.environ.pop("MB_RF_" + k, None)
        for k, v in kv.items():
            assert k in KNOBS, k
            os.environ["MB_RF_" + k] = v

        def run():
            torch.manual_seed(7)
            return m.sample(z, temperature=1.0, text_cfg=3.0, image_cfg=1.1 if cfg_rows == 3 else 1.0, groups=B // cfg_rows)

        for _ in range(2):
            out = run()
        ts = []
        for _ in range(7):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = run()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        if ref is None:
            ref = out.clone()
        buf = torch.zeros((16,), dtype=torch.int64, device=dev)
        lib.mb_rf_set_debug(buf.data_ptr())
        run()
        torch.cuda.synchronize()
        lib.mb_rf_set_debug(None)
        ph = buf.tolist()
        r = {"rows": B, "knobs": setting, "ms": round(ts[len(ts) // 2], 4), "ms_min": round(ts[0], 4),
             "gbs": round(alg / ts[len(ts) // 2] / 1e6, 1), "equal_to_first": bool(torch.equal(out, ref)),
             "phases_us_first_cta": {n: round(v / 1e3) for n, v in zip(names, ph[:7])},
             "phases_us_last_cta": {n: round(v / 1e3) for n, v in zip(names, ph[8:15])}}
        res.append(r)
        print(json.dumps(r), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
with open(os.environ.get("SWEEP_OUT", "gpurun_out/sweep_rf_l2.json"), "w") as f:
    json.dump(res, f, indent=1)
