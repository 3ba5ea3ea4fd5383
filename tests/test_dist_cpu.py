"""world_size-2 gloo check of bench.py's multi-rank plumbing (one process per GPU on the real box; here CPU): the
process-group setup from the torchrun environment, the barrier, and the max-over-ranks reduction of the timed duration.
MingTok batches shard over independent images with no data-path collective, so this is all the N > 1 logic there is."""
import os
import socket
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys, argparse
    sys.path.insert(0, %r)
    import torch, bench
    args = argparse.Namespace(gpus=2, steps=1, warmup=0, impl="ours")
    world, rank, local = bench._dist_setup(args)
    assert world == 2 and rank == int(os.environ["RANK"])
    bench._barrier(world)
    m = bench._max_over_ranks(10.0 + rank, world, torch.device("cpu"))
    assert m == 11.0, m
    # each rank runs its own edit round: different prompt / image per rank (bench.run_ours uses seed 100 * rank + i)
    from ming_univision_b200 import synthetic
    ids, um, tm, img = bench.round_inputs(synthetic.LLM_CONFIG, 100 * rank)
    assert ids.shape == (1, bench.PROMPT_LEN) and um.shape == tm.shape == (1, bench.PROMPT_LEN + 1)
    assert int((ids == synthetic.LLM_CONFIG["image_patch_token"]).sum()) == bench.N_ENC and img.dtype == torch.uint8
    assert int(tm.sum()) > 0 and not torch.equal(um, tm)          # -> 3 CFG rows (modeling_bailing_moe.py:1867-1891)
    import torch.distributed as dist
    t = img.flatten()[:64].float().clone()
    gathered = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(gathered, t)
    assert not torch.equal(gathered[0], gathered[1])
    dist.destroy_process_group()
    print("rank", rank, "ok")
""") % ROOT


EP_WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, %r)
    import torch, torch.distributed as dist
    from ming_univision_b200.ep import ExpertParallelAllToAll, token_slice
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    a2a = ExpertParallelAllToAll(None)
    assert (a2a.rank, a2a.size) == (rank, world)
    E, k, D, T = 6, 2, 8, 11                      # 3 experts per rank; uneven token slices (6 + 5)
    E_local = E // world
    g = torch.Generator().manual_seed(5)          # same stream on both ranks -> identical global problem
    x = torch.randn((T, D), generator=g)
    idx = torch.stack([torch.randperm(E, generator=g)[:k] for _ in range(T)]).to(torch.int32)
    w = torch.rand((T, k), generator=g)
    t0, t1, Tc = token_slice(T, world, rank)
    assert Tc == 6 and (t0, t1) == ((0, 6) if rank == 0 else (6, 11))
    xl, il, wl = x[t0:t1], idx[t0:t1], w[t0:t1]
    n = (t1 - t0) * k
    # send order = by destination rank (what mb_moe_plan with granule 1 produces on the GPU); torch stand-in here
    dest = (il.reshape(-1) // E_local).long()
    order = torch.argsort(dest, stable=True)                  # send row r holds pair order[r]
    pair_row = torch.empty(n, dtype=torch.long); pair_row[order] = torch.arange(n)
    counts = torch.bincount(dest, minlength=world).to(torch.int32)
    send_rows = xl[(order // k)]
    send_ids = il.reshape(-1)[order].contiguous()
    ss, rs = a2a.exchange_counts(counts)
    assert ss == counts.tolist() and sum(rs) >= 0
    recv_rows, recv_ids = a2a.dispatch(send_rows, send_ids, ss, rs)
    assert recv_rows.shape == (sum(rs), D)
    assert bool(((recv_ids // E_local) == rank).all()), "a row arrived at a rank that does not own its expert"
    # stand-in "expert": row * (expert id + 1)
    out = recv_rows * (recv_ids.float() + 1).unsqueeze(1)
    back = a2a.combine(out, ss, rs)
    assert back.shape == (n, D)
    per_pair = back[pair_row].view(t1 - t0, k, D)
    y_loc = torch.zeros((Tc, D)); y_loc[: t1 - t0] = (per_pair * wl.unsqueeze(-1)).sum(1)
    y = a2a.all_gather_rows(y_loc)[:T]
    ref = ((x.unsqueeze(1) * (idx.float() + 1).unsqueeze(-1)) * w.unsqueeze(-1)).sum(1)
    assert torch.allclose(y, ref, atol=1e-5), (y - ref).abs().max()
    # an empty sender: rank 1 sends nothing at all
    z = torch.zeros(world, dtype=torch.int32)
    c2 = counts if rank == 0 else z
    ss2, rs2 = a2a.exchange_counts(c2)
    r2, i2 = a2a.dispatch(send_rows if rank == 0 else send_rows[:0], send_ids if rank == 0 else send_ids[:0], ss2, rs2)
    b2 = a2a.combine(r2 * 2, ss2, rs2)
    if rank == 0:
        assert torch.allclose(b2, send_rows * 2)
    else:
        assert b2.shape[0] == 0
    dist.destroy_process_group()
    print("rank", rank, "ok")
""") % ROOT


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_two_rank_gloo_plumbing(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    port = _free_port()
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), CUDA_VISIBLE_DEVICES="")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for rank, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, f"rank {rank} failed:\n{o[-2000:]}"
        assert f"rank {rank} ok" in o


def test_expert_parallel_all_to_all_gloo(tmp_path):
    """ExpertParallelAllToAll (dispatch / combine / all-gather of the token- and expert-sharded MoE block) on two gloo
    ranks with CPU tensors and a stand-in expert function: the exchange must reproduce the unsharded weighted sum."""
    script = tmp_path / "ep_worker.py"
    script.write_text(EP_WORKER)
    port = _free_port()
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), CUDA_VISIBLE_DEVICES="")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for rank, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, f"rank {rank} failed:\n{o[-2000:]}"
        assert f"rank {rank} ok" in o


def test_reference_arm_other_ranks_are_silent():
    """`bench.py --impl reference` under torchrun: rank 0 alone prints the CPU-arm line; other ranks exit 0 silently."""
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1", CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], env=env,
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""
