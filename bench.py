#!/usr/bin/env python
"""Headline benchmark: continuous visual tokens/sec through the MingTok hot path (BASELINE.json configs[1]:
"MingTok ViT enc+dec batch=64 256x256 bf16, 1xB200").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A step = one `MingTok.forward_enc_dec` pass (low-level encoder -> causal semantic decoder -> pixel decoder) over one
batch of 64 synthetic 256x256 images = 4096 continuous latent tokens per rank.  Images are independent units, so
N ranks each process their own batch with replicated weights and NO data-path collective ("scaling": "weak");
the timed region is bracketed by barrier + synchronize and the reported time is the max over ranks.

  value    tokens/s with the step's inputs already resident in HBM (CUDA-event timed)
  e2e      the same metric through the public API with HOST buffers: pinned bf16 images are copied host->device and
           the reconstructed images device->host inside the timed region, every step (double-buffered on two copy
           streams, so step i's compute overlaps the upload of step i+1 and the download of step i-1)
  roofline dominant kernel (the tcgen05 GEMM): algorithmic FLOPs of every GEMM launch / CUDA-event duration of
           those launches, measured in an instrumented pass of the same step right after the timed region
  cpu_baseline / --impl reference: the CPU restatement of the reference (oracle/, kind "port" — the reference is
           pure Python/PyTorch and /root/reference does not exist on the GPU box) on the box's host cores, on a
           bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "continuous visual tokens/sec"
UNIT = "tokens/s"
BATCH, SIZE = 64, 256
TOKENS_PER_IMAGE = (SIZE // 32) ** 2
WORKLOAD = "MingTok ViT enc+dec batch=64 256x256 bf16 (BASELINE configs[1]); synthetic weights, synthetic images"


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"bf16_tflops": d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1450.0)),
                "hbm_gbs": d.get("hbm_gbs", 6489.0), "source": "measured (MEASURED_PEAKS.json, sustained bf16)"}
    return {"bf16_tflops": 1400.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """Samples SM clocks / throttle reasons with nvidia-smi while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.lines: list[str] = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def _dist_setup(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        backend = "nccl" if torch.cuda.is_available() else "gloo"
        dist.init_process_group(backend=backend, rank=rank, world_size=world,
                                device_id=torch.device("cuda", local) if backend == "nccl" else None)
    return world, rank, local


def _barrier(world):
    if world > 1:
        import torch.distributed as dist

        dist.barrier()


def _max_over_ranks(x: float, world: int, device) -> float:
    if world == 1:
        return x
    import torch.distributed as dist

    t = torch.tensor([x], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


# ---------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port on the host cores
# ---------------------------------------------------------------------------------------------------------------
_CPU_THREADS = None


def _best_cpu_threads(probe) -> int:
    """torch's CPU matmuls do not scale to every hardware thread of a 128-way host (measured: 16 threads are 4x
    faster than 128 on the B200 box), so the CPU arm uses the thread count that is fastest on a one-image probe."""
    global _CPU_THREADS
    if _CPU_THREADS is None:
        avail = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        best = (float("inf"), 1)
        with torch.no_grad():
            for th in sorted({t for t in (8, 16, 32, 64, avail) if t <= avail}):
                torch.set_num_threads(th)
                probe()
                t0 = time.perf_counter()
                probe()
                best = min(best, (time.perf_counter() - t0, th))
        _CPU_THREADS = best[1]
    return _CPU_THREADS


def cpu_oracle_sample(n_images: int, reps: int, seed: int = 1234):
    """Times the fp32 CPU restatement of the reference (oracle/mingtok_oracle.py) on `n_images` images of the
    workload; returns (tokens/s best-of-reps, cores, recon of the sample, images)."""
    from ming_univision_b200 import synthetic
    from oracle import mingtok_oracle as O

    cfg = synthetic.MINGTOK_CONFIG
    sd = synthetic.mingtok_state_dict(cfg, 0)
    img = synthetic.synthetic_images(BATCH, SIZE, seed=seed)[:n_images]
    cores = _best_cpu_threads(lambda: O.mingtok_forward_enc_dec(sd, img[:1], cfg))
    torch.set_num_threads(cores)
    best, recon = float("inf"), None
    with torch.no_grad():
        for _ in range(reps):
            t0 = time.perf_counter()
            recon = O.mingtok_forward_enc_dec(sd, img, cfg)
            best = min(best, time.perf_counter() - t0)
    return n_images * TOKENS_PER_IMAGE / best, cores, recon, img


def run_reference_arm(args, world, rank):
    """`--impl reference`: the reference's own CPU implementation of the path.  The reference is pure Python/PyTorch
    and cannot travel to the GPU box, so this is the oracle port (kind "port"), all host threads, bounded sample."""
    if rank != 0:
        return
    n_img = 4
    vals = []
    for _ in range(args.warmup if args.warmup < 1 else 1):
        cpu_oracle_sample(n_img, 1)
    for _ in range(max(1, min(args.steps, 3))):
        v, cores, _, _ = cpu_oracle_sample(n_img, 1)
        vals.append(v)
    value = statistics.median(vals)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * n_img * TOKENS_PER_IMAGE / value,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "sample": f"{n_img} of the {BATCH} images per step"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{n_img} images (256 latent tokens) per step, fp32, torch CPU, "
                                       f"{max(1, min(args.steps, 3))} timed steps, median"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------------
def run_ours(args, world, rank, local):
    from ming_univision_b200 import _lib, ops, synthetic
    from ming_univision_b200.mingtok import MingTok, MingTokConfig

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    _lib.require_device()
    cfg = synthetic.MINGTOK_CONFIG
    sd = synthetic.mingtok_state_dict(cfg, 0)
    with torch.device(dev):
        model = MingTok(MingTokConfig(**cfg))
    model.load_state_dict({k: v.to(dev) for k, v in sd.items()}, strict=True)
    model = model.to(torch.bfloat16)
    del sd

    # every rank gets its own images (seed offset by rank); several distinct batches are cycled so no step re-reads
    # the previous step's activations from L2 (weights alone, 1.4 GB, already exceed the 126 MB L2)
    n_bufs = 3
    host = [synthetic.synthetic_images(BATCH, SIZE, seed=1234 + 17 * rank + i).to(torch.bfloat16).pin_memory()
            for i in range(n_bufs)]
    dev_in = [h.to(dev) for h in host]

    def step_resident(i):
        return model.forward_enc_dec(dev_in[i % n_bufs])

    # end-to-end step through the public API with HOST buffers: every step copies its own batch host -> device and its
    # reconstruction device -> host (pinned memory).  The copies run on two copy streams (one per DMA direction) so that
    # step i's compute overlaps the upload of step i + 1 and the download of step i - 1 — ordinary double buffering of a
    # serving loop; every byte still moves inside the timed region.
    h2d_stream, d2h_stream = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    host_outs = [torch.empty((BATCH, 3, SIZE, SIZE), dtype=torch.bfloat16).pin_memory() for _ in range(2)]
    dev_x = [torch.empty((BATCH, 3, SIZE, SIZE), dtype=torch.bfloat16, device=dev) for _ in range(2)]

    def run_e2e(n_steps):
        main = torch.cuda.current_stream(dev)
        up = [torch.cuda.Event() for _ in range(n_steps)]
        done = [torch.cuda.Event() for _ in range(n_steps)]
        down = [torch.cuda.Event() for _ in range(n_steps)]

        def upload(i):
            with torch.cuda.stream(h2d_stream):
                if i >= 2:
                    h2d_stream.wait_event(done[i - 2])  # the device buffer is free once step i - 2 has consumed it
                dev_x[i % 2].copy_(host[i % n_bufs], non_blocking=True)
                up[i].record(h2d_stream)

        h2d_stream.wait_stream(main)
        d2h_stream.wait_stream(main)
        upload(0)
        y = None
        for i in range(n_steps):
            if i + 1 < n_steps:
                upload(i + 1)
            main.wait_event(up[i])
            y = model.forward_enc_dec(dev_x[i % 2])
            done[i].record(main)
            with torch.cuda.stream(d2h_stream):
                d2h_stream.wait_event(done[i])
                if i >= 2:
                    d2h_stream.wait_event(down[i - 2])
                host_outs[i % 2].copy_(y, non_blocking=True)
                down[i].record(d2h_stream)
            y.record_stream(d2h_stream)
        main.wait_stream(d2h_stream)
        main.wait_stream(h2d_stream)
        return y

    for i in range(args.warmup):
        step_resident(i)
    run_e2e(args.warmup)
    torch.cuda.synchronize()

    # ---- timed region 1: inputs resident in HBM
    sampler = ClockSampler(local) if rank == 0 else None
    _barrier(world)
    torch.cuda.synchronize()
    if sampler:
        sampler.start()
    launches0 = _lib.launch_count()
    ncu_range = os.environ.get("MB_NCU_RANGE") == "1"  # tools/profile.sh: ncu --profile-from-start off
    if ncu_range:
        torch.cuda.profiler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step_resident(i)
    e1.record()
    torch.cuda.synchronize()
    if ncu_range:
        torch.cuda.profiler.stop()
    launches = _lib.launch_count() - launches0
    _barrier(world)
    ms_res = _max_over_ranks(e0.elapsed_time(e1), world, dev)

    # ---- timed region 2: end to end through the public API with host buffers
    _barrier(world)
    torch.cuda.synchronize()
    e0.record()
    run_e2e(args.steps)
    e1.record()
    torch.cuda.synchronize()
    _barrier(world)
    ms_e2e = _max_over_ranks(e0.elapsed_time(e1), world, dev)
    clocks = sampler.stop() if sampler else None

    tokens_per_step = BATCH * TOKENS_PER_IMAGE * world
    value = tokens_per_step * args.steps / (ms_res / 1e3)
    e2e_value = tokens_per_step * args.steps / (ms_e2e / 1e3)

    if rank != 0:
        return

    # ---- roofline of the dominant kernel (the tcgen05 GEMM).  Two live measurements over the launches of one step:
    # (a) replay: every GEMM launch of the step is recorded (same buffers / shapes / epilogues) and the whole list is
    #     re-issued back to back on the launching stream between two CUDA events -> average launch duration with the PDL
    #     chain intact;  (b) instrumented: an event pair around every launch inside a real step (serialises the chain and
    #     adds ~2 us per launch, so it reads low; kept for transparency).
    ops.PROFILE, ops.REPLAY = [], []
    step_resident(0)
    torch.cuda.synchronize()
    prof, ops.PROFILE = ops.PROFILE, None
    replay, ops.REPLAY = ops.REPLAY, None
    gemm_ms_instr = sum(s.elapsed_time(e) for (_, _, s, e) in prof)
    gemm_flops = sum(f for (_, f, _, _) in prof)
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for fn, _, _ in replay:  # warm-up of the replay list
        fn()
    torch.cuda.synchronize()
    reps = 3
    t0.record()
    for _ in range(reps):
        for fn, _, _ in replay:
            fn()
    t1.record()
    torch.cuda.synchronize()
    gemm_ms = t0.elapsed_time(t1) / reps
    del replay
    t0.record()
    step_resident(0)
    t1.record()
    torch.cuda.synchronize()
    step_ms = t0.elapsed_time(t1)
    peaks = _peaks()
    traffic, traffic_src = None, None
    try:  # DRAM traffic per launch from the committed ncu --set full capture of the pixel-decoder GEMMs
        with open(os.path.join(ROOT, "profiles", "gemm_traffic.json")) as f:
            tj = json.load(f)
        traffic, traffic_src = tj["mean_dram_bytes_per_launch"], tj["source"]
    except (OSError, KeyError, ValueError):
        pass
    achieved = gemm_flops / (gemm_ms / 1e3) / 1e12
    roofline = {"bound": "tensor", "kernel": "mb::gemm_bf16_kernel (tcgen05.mma + TMA)", "achieved": achieved,
                "peak": peaks["bf16_tflops"], "unit": "TFLOP/s", "frac": achieved / peaks["bf16_tflops"],
                "peak_source": peaks["source"], "traffic": traffic, "traffic_unit": "bytes per launch (dram read + write)",
                "traffic_source": traffic_src, "gemm_launches_per_step": len(prof),
                "gemm_ms_per_step": gemm_ms, "gemm_share_of_step": gemm_ms / step_ms,
                "avg_launch_us": 1e3 * gemm_ms / max(len(prof), 1), "timing": "replay of the step's GEMM launches, CUDA events",
                "achieved_instrumented": gemm_flops / (gemm_ms_instr / 1e3) / 1e12, "gemm_ms_instrumented": gemm_ms_instr,
                "gemm_flops_per_step": gemm_flops,
                "step_flops_algorithmic": 213.0e9 * BATCH,
                "step_tflops_algorithmic": 213.0e9 * BATCH * world / (ms_res / args.steps / 1e3) / 1e12}

    # ---- CPU baseline (oracle port) on a bounded sample of the same workload + parity of that sample
    n_cpu = 4
    cpu_val, cores, ref_recon, img = cpu_oracle_sample(n_cpu, 2)
    ours = model.forward_enc_dec(img.to(dev).to(torch.bfloat16)).float().cpu()
    mse_o = float(((ours - img) ** 2).mean())
    mse_r = float(((ref_recon - img) ** 2).mean())
    import math
    parity = {"n_images": n_cpu,
              "rel_l2_recon_vs_fp32_oracle": float((ours - ref_recon).norm() / ref_recon.norm()),
              "psnr_ours_vs_oracle_db": 10 * math.log10(4.0 / max(float(((ours - ref_recon) ** 2).mean()), 1e-30)),
              "dpsnr_vs_input_db": abs(10 * math.log10(4.0 / mse_o) - 10 * math.log10(4.0 / mse_r))}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_res / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "global_batch": BATCH * world, "tokens_per_step": tokens_per_step,
                       "parallelism": f"dp{world} (independent image batches, replicated weights, no collective)",
                       "l2_policy": "3 distinct input batches cycled; weights (1.4 GB) + activations exceed L2"},
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": BATCH * 3 * SIZE * SIZE * 2, "d2h_bytes_per_step": BATCH * 3 * SIZE * SIZE * 2},
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline,
            "cpu_baseline": {"value": cpu_val, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{n_cpu} images of the batch (fp32 torch CPU restatement, best of 2)"},
            "parity": parity}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":  # CPU arm: rank 0 alone works, no process group needed
        run_reference_arm(args, int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")))
        return
    world, rank, local = _dist_setup(args)
    try:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device (the ming_univision_b200 path has no CPU fallback)")
        run_ours(args, world, rank, local)
    finally:
        if world > 1:
            import torch.distributed as dist

            if dist.is_initialized():
                dist.destroy_process_group()


if __name__ == "__main__":
    main()
