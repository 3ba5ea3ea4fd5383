"""Expert parallelism over 2 GPUs (one process per GPU, NCCL): the MoE block with its routed experts sharded over the
ranks must reproduce the single-GPU block (same router decisions; fp32 partial sums reduced in a different order, so
equality is to bf16 rounding), and every rank must end with the same output — both for the replicated-token all-reduce
mode (decode) and for the token-sharded all-to-all dispatch / combine mode (prefill).  Needs >= 2 GPUs
(`gpurun --gpus 2`)."""
import os
import socket
import subprocess
import sys
import textwrap

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, %r)
    import torch, torch.distributed as dist
    from ming_univision_b200 import synthetic
    from ming_univision_b200.modeling_bailing_moe import BailingMoeConfig, BailingMoeSparseMoeBlock
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    cfg = dict(synthetic.LLM_TINY_CONFIG, num_experts=16, num_experts_per_tok=6, moe_intermediate_size=96,
               num_shared_experts=2, hidden_size=256)
    sd_all = synthetic.llm_state_dict(dict(cfg, num_hidden_layers=1), None, None, seed=3)
    pre = "model.layers.0.mlp."
    sd = {k[len(pre):]: v.to(dev) for k, v in sd_all.items() if k.startswith(pre)}
    def build():
        with torch.device(dev):
            b = BailingMoeSparseMoeBlock(BailingMoeConfig(**cfg))
        b.load_state_dict(sd, strict=True)
        return b.to(torch.bfloat16)
    single, ep = build(), build()
    ep.set_expert_parallel(dist.group.WORLD, rank, world)
    g = torch.Generator().manual_seed(7)
    for T in (1, 3, 40):
        x = torch.randn((1, T, cfg["hidden_size"]), generator=g).to(dev).to(torch.bfloat16)
        res = torch.randn((T, cfg["hidden_size"]), generator=g).to(dev).to(torch.bfloat16)
        y1, _, i1 = single._run(x.view(T, -1), res, None)
        y2, _, i2 = ep._run(x.view(T, -1), res, None)
        assert torch.equal(i1, i2)
        err = ((y1.float() - y2.float()).norm() / y1.float().norm()).item()
        assert err < 5e-3, err
        gathered = [torch.empty_like(y2) for _ in range(world)]
        dist.all_gather(gathered, y2)
        assert all(torch.equal(gathered[0], t) for t in gathered), "ranks disagree"
        print("rank", rank, "T", T, "rel err vs single GPU", err, flush=True)
    # fused exchange over NVLink peer memory (no NCCL in the combine): decode-sized inputs, several calls in a row so
    # that both parities and the epoch protocol are exercised; larger inputs fall back to the all-reduce
    from ming_univision_b200.ep import PeerExchange
    peer_blk = build()
    px = PeerExchange(dist.group.WORLD, cfg["hidden_size"], dev)
    peer_blk.set_expert_parallel(dist.group.WORLD, rank, world, mode="peer", peer=px)
    for it, T in enumerate((1, 3, 2, 8, 3, 40)):
        x = torch.randn((1, T, cfg["hidden_size"]), generator=g).to(dev).to(torch.bfloat16)
        res = torch.randn((T, cfg["hidden_size"]), generator=g).to(dev).to(torch.bfloat16)
        y1, _, _ = single._run(x.view(T, -1), res, None)
        y2, _, _ = peer_blk._run(x.view(T, -1), res, None)
        err = ((y1.float() - y2.float()).norm() / y1.float().norm()).item()
        assert err < 5e-3, err
        gathered = [torch.empty_like(y2) for _ in range(world)]
        dist.all_gather(gathered, y2)
        assert all(torch.equal(gathered[0], t) for t in gathered), "ranks disagree (peer mode)"
        print("rank", rank, "T", T, "peer-memory rel err vs single GPU", err, flush=True)
    # "dispatch" mode (data parallel x expert parallel): every rank runs its OWN rows; dispatch and combine both go
    # through peer memory inside the kernels of csrc/ep.cu.  Each rank's output must match the unsharded block on ITS rows.
    from ming_univision_b200.ep import PeerDispatch
    dblk = build()
    pd = PeerDispatch(dist.group.WORLD, cfg["hidden_size"], cfg["num_experts_per_tok"], 64, dev)
    dblk.set_expert_parallel(dist.group.WORLD, rank, world, mode="dispatch", peer=pd)
    gr = torch.Generator().manual_seed(100 + rank)   # different rows on every rank
    for T in (1, 3, 2, 8, 3, 40, 150):
        x = torch.randn((1, T, cfg["hidden_size"]), generator=gr).to(dev).to(torch.bfloat16)
        res = torch.randn((T, cfg["hidden_size"]), generator=gr).to(dev).to(torch.bfloat16)
        y1, _, i1 = single._run(x.view(T, -1), res, None)
        y2, _, i2 = dblk._run(x.view(T, -1), res, None)
        torch.cuda.synchronize(); pd.check()
        assert torch.equal(i1, i2)
        err = ((y1.float() - y2.float()).norm() / y1.float().norm()).item()
        assert err < 5e-3, err
        print("rank", rank, "T", T, "dispatch-mode rel err vs single GPU", err, flush=True)
    # token- AND expert-sharded block: all-to-all dispatch / combine + all-gather (prefill-sized inputs; T = 7 stays
    # on the all-reduce path, T = 41 gives uneven token slices 21 + 20)
    a2a = build()
    a2a.set_expert_parallel(dist.group.WORLD, rank, world, mode="alltoall")
    os.environ["MB_MOE_GROUPED"] = "1"
    for T in (7, 41, 300):
        x = torch.randn((1, T, cfg["hidden_size"]), generator=g).to(dev).to(torch.bfloat16)
        res = torch.randn((T, cfg["hidden_size"]), generator=g).to(dev).to(torch.bfloat16)
        im = (torch.arange(T) %% 3 == 0).to(dev)
        y1, _, _ = single._run(x.view(T, -1), res, im)
        y2, _, _ = a2a._run(x.view(T, -1), res, im)
        assert y2.shape == y1.shape
        err = ((y1.float() - y2.float()).norm() / y1.float().norm()).item()
        assert err < 5e-3, err
        gathered = [torch.empty_like(y2) for _ in range(world)]
        dist.all_gather(gathered, y2.contiguous())
        assert all(torch.equal(gathered[0], t) for t in gathered), "ranks disagree"
        print("rank", rank, "T", T, "all-to-all rel err vs single GPU", err, flush=True)
    dist.destroy_process_group()
    print("rank", rank, "ok")
""") % ROOT


def test_expert_parallel_two_ranks(tmp_path):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    script = tmp_path / "ep_worker.py"
    script.write_text(WORKER)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=600)[0] for p in procs]
    for rank, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, f"rank {rank} failed:\n{o[-3000:]}"
        assert f"rank {rank} ok" in o
