"""Expert-parallel decode step of ONE full-size Bailing-MoE MoE layer (64 experts top-6 + 2 shared, D = 2048) under
torchrun: experts sharded over the ranks, tokens (B = 2 CFG rows) replicated, fp32 partial sums all-reduced with NCCL.
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_ep.py
Prints per-layer latency (CUDA events, max over ranks) for EP = N next to the unsharded layer on rank 0."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ming_univision_b200 import synthetic  # noqa: E402
from ming_univision_b200.modeling_bailing_moe import BailingMoeConfig, BailingMoeSparseMoeBlock  # noqa: E402

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
if world > 1:
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
cfg = BailingMoeConfig(**synthetic.LLM_CONFIG)
torch.manual_seed(0)
torch.set_default_dtype(torch.bfloat16)
with torch.device(dev):
    blk = BailingMoeSparseMoeBlock(cfg)
torch.set_default_dtype(torch.float32)
with torch.no_grad():
    for p in blk.parameters():
        p.copy_(torch.randn(p.shape, device=dev, dtype=torch.float32) / p.shape[-1] ** 0.5)
x = torch.randn((2, 2048), device=dev).to(torch.bfloat16)
res = torch.randn((2, 2048), device=dev).to(torch.bfloat16)
if world > 1:
    dist.broadcast(x, 0)
    dist.broadcast(res, 0)
    for p in blk.parameters():
        dist.broadcast(p.data, 0)


def timeit(fn, iters=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    t = torch.tensor([s.elapsed_time(e) / iters], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


y_single, _, _ = blk._run(x, res, None)
ms_single = timeit(lambda: blk._run(x, res, None))
out = {"world": world, "ms_layer_unsharded": round(ms_single, 4)}
if world > 1:
    blk.set_expert_parallel(dist.group.WORLD, rank, world)
    y_ep, _, _ = blk._run(x, res, None)
    out["rel_err_vs_unsharded"] = float(((y_ep.float() - y_single.float()).norm() / y_single.float().norm()).item())
    out["ms_layer_ep"] = round(timeit(lambda: blk._run(x, res, None)), 4)
    from ming_univision_b200.ep import PeerExchange  # noqa: E402
    px = PeerExchange(dist.group.WORLD, 2048, dev)
    blk.set_expert_parallel(dist.group.WORLD, rank, world, mode="peer", peer=px)
    y_px, _, _ = blk._run(x, res, None)
    out["rel_err_peer_vs_unsharded"] = float(((y_px.float() - y_single.float()).norm() / y_single.float().norm()).item())
    out["ms_layer_ep_peer"] = round(timeit(lambda: blk._run(x, res, None)), 4)
# ---- "dispatch" mode (data parallel x expert parallel, csrc/ep.cu): every rank has its OWN rows; the MoE part of a token
# step (28 consecutive layer calls) captured in ONE CUDA graph on both sides, so host launch time is out of the picture
def graph_ms(fn, layers=28, reps=10):
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
    torch.cuda.current_stream().wait_stream(side)
    if world > 1:
        dist.barrier()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(layers):
            fn()
    g.replay()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        g.replay()
    e.record()
    torch.cuda.synchronize()
    t = torch.tensor([s.elapsed_time(e) / reps / layers], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


if world > 1:
    blk.set_expert_parallel(None, 0, 1)
for R in (3, 6):
    xr = torch.randn((R, 2048), generator=torch.Generator(device=dev).manual_seed(100 + rank), device=dev).to(torch.bfloat16)
    rr = torch.randn((R, 2048), generator=torch.Generator(device=dev).manual_seed(200 + rank), device=dev).to(torch.bfloat16)
    out[f"ms_layer_graph_unsharded_rows{R}"] = round(graph_ms(lambda: blk._run(xr, rr, None)), 4)
if world > 1:
    from ming_univision_b200.ep import PeerDispatch  # noqa: E402
    pd = PeerDispatch(dist.group.WORLD, 2048, cfg.num_experts_per_tok, 64, dev)
    blk.set_expert_parallel(dist.group.WORLD, rank, world, mode="dispatch", peer=pd)
    for R in (3, 6):
        xr = torch.randn((R, 2048), generator=torch.Generator(device=dev).manual_seed(100 + rank), device=dev).to(torch.bfloat16)
        rr = torch.randn((R, 2048), generator=torch.Generator(device=dev).manual_seed(200 + rank), device=dev).to(torch.bfloat16)
        out[f"ms_layer_graph_dispatch_rows{R}"] = round(graph_ms(lambda: blk._run(xr, rr, None)), 4)
        torch.cuda.synchronize()
        pd.check()
    blk.set_expert_parallel(None, 0, 1)

# ---- prefill-sized input (BASELINE configs[2] shape: 1552 tokens): unsharded grouped tcgen05 GEMMs vs expert-parallel
# all-reduce (tokens replicated) vs token-sharded all-to-all dispatch / combine + all-gather
T = int(os.environ.get("EP_PREFILL_TOKENS", "1552"))
xp = torch.randn((T, 2048), device=dev).to(torch.bfloat16)
rp = torch.randn((T, 2048), device=dev).to(torch.bfloat16)
if world > 1:
    dist.broadcast(xp, 0)
    dist.broadcast(rp, 0)
    blk.set_expert_parallel(None, 0, 1)
yp_single, _, _ = blk._run(xp, rp, None)
out["prefill_tokens"] = T
out["ms_prefill_layer_unsharded"] = round(timeit(lambda: blk._run(xp, rp, None), iters=20), 4)
if world > 1:
    for mode in ("allreduce", "alltoall"):
        blk.set_expert_parallel(dist.group.WORLD, rank, world, mode=mode)
        yp, _, _ = blk._run(xp, rp, None)
        out[f"prefill_rel_err_{mode}"] = float(((yp.float() - yp_single.float()).norm() / yp_single.float().norm()).item())
        out[f"ms_prefill_layer_ep_{mode}"] = round(timeit(lambda: blk._run(xp, rp, None), iters=20), 4)
if rank == 0:
    print(json.dumps(out), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    with open(f"gpurun_out/bench_ep_{world}.json", "w") as f:
        json.dump(out, f, indent=1)
if world > 1:
    dist.destroy_process_group()
