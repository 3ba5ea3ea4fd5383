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
    # each rank draws its own images: different seeds per rank (bench.run_ours uses 1234 + 17 * rank + i)
    from ming_univision_b200 import synthetic
    a = synthetic.synthetic_images(1, 64, seed=1234 + 17 * rank)
    import torch.distributed as dist
    t = a.flatten()[:8].clone()
    gathered = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(gathered, t)
    assert not torch.equal(gathered[0], gathered[1])
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


def test_reference_arm_other_ranks_are_silent():
    """`bench.py --impl reference` under torchrun: rank 0 alone prints the CPU-arm line; other ranks exit 0 silently."""
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1", CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], env=env,
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""
