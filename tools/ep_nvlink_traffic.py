"""NVLink evidence for the expert-parallel peer-memory exchange (csrc/ep.cu): NVLink data counters of this rank's GPU
before / after N MoE layers in "dispatch" mode, next to the bytes the protocol must move.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/ep_nvlink_traffic.py
Counters: NVML field values NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_TX / _RX (KiB, all links), with `nvidia-smi nvlink -gt d`
as a cross-check.  Per layer and rank the kernels store to each of the G - 1 peers: T rows x D bf16 + T k (i32 + f32)
(dispatch) and T rows x D fp32 (this rank's partial sums of the PEER's rows) (combine)."""
import json
import os
import subprocess
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ming_univision_b200 import synthetic  # noqa: E402
from ming_univision_b200.ep import PeerDispatch  # noqa: E402
from ming_univision_b200.modeling_bailing_moe import BailingMoeConfig, BailingMoeSparseMoeBlock  # noqa: E402

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
cfg = BailingMoeConfig(**synthetic.LLM_CONFIG)
torch.set_default_dtype(torch.bfloat16)
with torch.device(dev):
    blk = BailingMoeSparseMoeBlock(cfg)
torch.set_default_dtype(torch.float32)
torch.manual_seed(0)
with torch.no_grad():
    for p in blk.parameters():
        p.copy_(torch.randn(p.shape, device=dev, dtype=torch.float32) / p.shape[-1] ** 0.5)
pd = PeerDispatch(dist.group.WORLD, 2048, cfg.num_experts_per_tok, 64, dev)
blk.set_expert_parallel(dist.group.WORLD, rank, world, mode="dispatch", peer=pd)
T, D, k = 6, 2048, cfg.num_experts_per_tok
x = torch.randn((T, D), device=dev).to(torch.bfloat16)
res = torch.randn((T, D), device=dev).to(torch.bfloat16)
blk._run(x, res, None)
torch.cuda.synchronize()


def counters():
    out = {}
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(rank)
        vals = pynvml.nvmlDeviceGetFieldValues(h, [138, 139])  # NVLINK_THROUGHPUT_DATA_TX / RX, KiB
        out["nvml_tx_kib"], out["nvml_rx_kib"] = int(vals[0].value.ullVal), int(vals[1].value.ullVal)
    except Exception as e:  # noqa: BLE001
        out["nvml_error"] = repr(e)[:200]
    try:
        r = subprocess.run(["nvidia-smi", "nvlink", "-gt", "d", "-i", str(rank)], capture_output=True, text=True, timeout=30)
        tx = rx = 0
        for ln in r.stdout.splitlines():
            if "Data Tx" in ln:
                tx += int(ln.split(":")[-1].strip().split()[0])
            if "Data Rx" in ln:
                rx += int(ln.split(":")[-1].strip().split()[0])
        out["smi_tx_kib"], out["smi_rx_kib"] = tx, rx
    except Exception as e:  # noqa: BLE001
        out["smi_error"] = repr(e)[:200]
    return out


dist.barrier()
c0 = counters()
N = int(os.environ.get("EP_LAYERS", "20000"))
for _ in range(N):
    blk._run(x, res, None)
torch.cuda.synchronize()
pd.check()
dist.barrier()
c1 = counters()
expected = N * (world - 1) * (T * D * 2 + T * k * 8 + T * D * 4)
out = {"world": world, "rank": rank, "layers": N, "rows_per_rank": T,
       "expected_tx_bytes_per_rank": expected, "expected_tx_kib": expected // 1024,
       "delta": {k2: c1[k2] - c0[k2] for k2 in c1 if k2 in c0 and isinstance(c1[k2], int)}, "before": c0, "after": c1}
print(json.dumps(out), flush=True)
if rank == 0:
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/ep_nvlink_traffic.json", "w") as f:
        json.dump(out, f, indent=1)
dist.destroy_process_group()
