#!/usr/bin/env python
"""Headline benchmark: continuous visual tokens/sec through enc + AR + RF + dec at 256x256 (BASELINE.json `metric`) on the
full-size synthetic Ming-UniVision 16B-A3B (28-layer Bailing-MoE, 64 experts top-6, default RF head, full MingTok).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A step = ONE in-context EDIT ROUND at 256x256 per GPU (reference: mingunivisioninfer.py:82-117 with for_edit=True ->
modeling_bailingmm.py:206-301 -> modeling_bailing_moe.py:1844-1965 -> diff_loss_rf_swiglu.py:103-181):
    enc   MingTok low-level encoder + causal semantic decoder on the 256x256 input image (64 visual tokens) + linear_proj
    AR    prefill of the 105-token prompt (text + the 64 image embeddings routed by image_gate), then 64 (+1 discarded,
          as in the reference) AR steps with B = 3 CFG rows (cond / uncond / text-uncond): 28-layer MoE step -> vis_head
    RF    per AR step: rectified-flow SwiGLU head, 16 Euler steps x 3 rows, CFG 3.0 / 1.1 -> latent [32]
          -> cached semantic-decoder step -> linear_proj -> next input embedding
    dec   pixel decoder on the 64 generated tokens -> 256x256 image
= 64 encoded + 64 generated continuous visual tokens ("(encoded + generated) / time"); the generated
tokens alone are reported next to it (`generated_tokens_per_s`).

  value     tokens/s with the step's inputs (normalised bf16 image, token ids) resident in HBM; CUDA-event timed
  e2e       the same through the public API from HOST buffers: pinned u8 image + ids host->device, on-device
            normalisation, the round, on-device u8 conversion, pinned u8 image device->host — all inside the timed region
  roofline  bound "hbm": the dominant kernel, mb::rf_sample_fused_kernel (the whole RF sampler of a token step as one
            persistent weight-streaming launch): algorithmic bytes per launch (the bf16 weights it must stream: steps x
            depth x (2H*W + W*H) x 2 = 28.99 GB) / the median CUDA-event duration of its launches on the launching
            stream, against the measured copy bandwidth in MEASURED_PEAKS.json; `traffic` from the committed ncu capture;
            `second_kernel`: the per-layer streaming GEMM (mb::gemv_bf16_kernel: LLM dense layers, semantic decoder, heads)
  stages    secondary: MingTok enc+dec batch=64 256x256 (BASELINE configs[1], last round's headline) with its tensor-core
            roofline, and the per-stage split of the round
  cpu_baseline / --impl reference: the CPU restatement of the reference (oracle/, kind "port": the reference is pure
            Python/PyTorch and /root/reference does not exist on the GPU box) on the box's host cores, on a BOUNDED
            sample of the same round — the LLM depth-reduced to 2 of its 28 layers at the true widths and
            2 of the 65 AR steps, every stage timed separately and extrapolated (flagged) to the full round.

Batched serving: `--images G` (default 2) edit rounds are generated TOGETHER on a GPU — G x 3 CFG rows share every pass
over the LLM / RF-head / semantic-decoder weights (the reference asserts one sequence, modeling_bailing_moe.py:1865; every
request's image is bit-identical to what it gets alone, tests/test_llm_wide_gpu.py).  tokens per step = G x 128; the
one-request-at-a-time number is reported as stages.single_request.

N > 1: every rank runs ITS OWN rounds (its own image / prompt) and the routed experts are SHARDED over the ranks
(expert parallelism, 64 / N experts per GPU): per MoE layer the rows are dispatched to the expert owners and the partial
sums combined back through NVLink peer memory inside the kernels (csrc/ep.cu; no NCCL call, the whole token step stays
one CUDA graph).  Per-GPU work is fixed -> "scaling": "weak"; value = all ranks' tokens / max-over-ranks time.
`ep_check`: every rank also generates ONE identical request; the image digests of all ranks must agree (the exchange is
deterministic and reduces in rank order) — the consistency check a multi-GPU box can run.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "continuous visual tokens/sec (enc+AR+RF+dec) at 256x256"
UNIT = "tokens/s"
SIZE = 256
N_GEN = (SIZE // 32) ** 2          # 64 generated visual tokens (num_image_tokens_for_gen for a 256x256 image)
N_ENC = (SIZE // 32) ** 2          # 64 encoded visual tokens of the input image
TOKENS_PER_STEP = N_ENC + N_GEN
CFG_ROWS = 3
N_TEXT_HEAD, N_TEXT_TAIL = 8, 32   # prompt = 8 text ids + 64 <imagePatch> + 32 text ids  (+ the <image> start token)
PROMPT_LEN = N_TEXT_HEAD + N_ENC + N_TEXT_TAIL
WORKLOAD = ("Ming-UniVision 16B-A3B in-context edit rounds at 256x256, each: MingTok enc (64 tokens) -> 104-token prefill -> "
            "64 AR steps x (28-layer MoE, B=3 CFG rows, RF head 16 Euler steps, semantic-decoder step) -> pixel decoder; "
            "synthetic weights, synthetic image / prompt")
MINGTOK_BATCH = 64


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"bf16_tflops": d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1450.0)),
                "hbm_gbs": d.get("hbm_gbs", 6489.0), "source": "measured (MEASURED_PEAKS.json)"}
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


def _dist_setup(args=None):
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
# the synthetic edit round: prompt, masks, image
# ---------------------------------------------------------------------------------------------------------------
def round_inputs(llm_cfg: dict, seed: int):
    """Token ids [1, 104] (text ids uniform in [0, 126000), the 64 image positions = <imagePatch>), the two CFG masks
    over prompt + <image> start token, and a u8 HWC image — what processing_bailingmm would hand over for an edit."""
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(0, 126000, (1, PROMPT_LEN), generator=g)
    ids[:, N_TEXT_HEAD:N_TEXT_HEAD + N_ENC] = llm_cfg["image_patch_token"]
    n = PROMPT_LEN + 1
    uncond = torch.ones((1, n), dtype=torch.int32)          # uncond row: drops text AND image, keeps the frame tokens
    uncond[:, 2:PROMPT_LEN - 3] = 0
    text_uncond = torch.ones((1, n), dtype=torch.int32)     # text-uncond row: keeps the image, drops the instruction
    text_uncond[:, N_TEXT_HEAD + N_ENC:PROMPT_LEN - 3] = 0
    from ming_univision_b200 import synthetic

    img = synthetic.synthetic_images(1, SIZE, seed=1234 + seed)                      # [-1, 1] fp32 NCHW
    img_u8 = ((img[0].permute(1, 2, 0) * 0.5 + 0.5) * 255).round().clamp(0, 255).to(torch.uint8).contiguous()
    return ids, uncond, text_uncond, img_u8


# ---------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port on the host cores (bounded sample, every stage timed, extrapolated to the full round)
# ---------------------------------------------------------------------------------------------------------------
def _pick_cpu_threads(probe) -> int:
    """torch's CPU matmuls do not scale to every hardware thread of a 128-way host (measured: 16 threads are 4x faster than
    128 on the B200 box), so the CPU arm uses the thread count that is fastest on a probe."""
    avail = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    best = (float("inf"), 1)
    with torch.no_grad():
        for th in sorted({t for t in (8, 16, 32, 64, avail) if t <= avail}):
            torch.set_num_threads(th)
            probe()
            t0 = time.perf_counter()
            probe()
            best = min(best, (time.perf_counter() - t0, th))
    return best[1]


class CpuReference:
    """fp32 restatement of the reference (oracle/) for one edit round, LLM depth-reduced to 2 layers at true widths."""
    LAYERS = 2
    AR_STEPS = 2          # timed AR steps per sample (of the round's 65)

    def __init__(self):
        from ming_univision_b200 import synthetic
        from oracle import bailing_oracle as L
        from oracle import mingtok_oracle as O

        self.L, self.O, self.syn = L, O, synthetic
        self.cfg = dict(synthetic.LLM_CONFIG, num_hidden_layers=self.LAYERS, num_image_tokens_for_gen=N_GEN)
        self.vh, self.tok_cfg = synthetic.VISHEAD_CONFIG, synthetic.MINGTOK_CONFIG
        t0 = time.perf_counter()
        torch.set_num_threads(min(32, os.cpu_count() or 8))
        self.sd = synthetic.llm_state_dict(self.cfg, self.vh, feature_dim=1024, seed=0)
        self.rf_sd = {k[len("diffloss."):]: v for k, v in self.sd.items() if k.startswith("diffloss.")}
        self.tok_sd = synthetic.mingtok_state_dict(self.tok_cfg, 0)
        self.build_s = time.perf_counter() - t0
        self.ids, self.um, self.tm, self.img_u8 = round_inputs(self.cfg, 0)
        z = torch.randn(CFG_ROWS, 3072)
        noise = torch.randn(1, 32)
        from oracle import rf_oracle as R

        self.cores = _pick_cpu_threads(lambda: R.sample(self.rf_sd, z, noise, 2, 1.0, 3.0, 1.1))
        torch.set_num_threads(self.cores)

    @torch.no_grad()
    def sample_step(self, keep: bool = False) -> dict:
        """One bounded sample of the round; returns the per-stage seconds (+ tensors for the parity check if `keep`)."""
        L, O, sd, cfg = self.L, self.O, self.sd, self.cfg
        t = {}
        img = (self.img_u8.float() / 255.0 - 0.5) / 0.5
        img = img.permute(2, 0, 1).unsqueeze(0).contiguous()
        c0 = time.perf_counter()
        enc = O.mingtok_forward(self.tok_sd, img, self.tok_cfg)
        vis = L.linear_proj(sd, enc["x_norm_patchtokens"])
        t["enc"] = time.perf_counter() - c0
        emb = sd["model.word_embeddings.weight"][self.ids].clone()
        image_mask = self.ids == cfg["image_patch_token"]
        emb[image_mask] = vis.reshape(-1, vis.shape[-1])
        caches = L.new_caches(cfg)
        c0 = time.perf_counter()
        h = L.model_forward(sd, cfg, emb, torch.ones(1, PROMPT_LEN, dtype=torch.long), None, caches, image_mask=image_mask)
        t["prefill_2layers"] = time.perf_counter() - c0
        start = sd["model.word_embeddings.weight"][torch.tensor([[cfg["image_start_token"]]])]
        sem_t, state_holder = [0.0], {}

        def latent_to_sem(latent, state):
            s0 = time.perf_counter()
            state = O.new_decoder_caches(self.tok_sd) if state is None else state
            r = O.mingtok_forward_feature_decoder(self.tok_sd, latent, self.tok_cfg, state)
            sem_t[0] += time.perf_counter() - s0
            return r, state

        g = torch.Generator().manual_seed(11)
        noises = [torch.randn(1, 32, generator=g) for _ in range(self.AR_STEPS + 1)]
        c0 = time.perf_counter()
        feats, lats, _ = L.generate_image(sd, cfg, self.rf_sd, int(self.vh["num_sampling_steps"]), start, caches,
                                          torch.ones(1, PROMPT_LEN + 1, dtype=torch.long), self.um.long(), self.tm.long(),
                                          latent_to_sem, lambda f: L.linear_proj(sd, f), noises, temperature=1.0,
                                          num_tokens=self.AR_STEPS - 1)
        t["ar_steps"] = time.perf_counter() - c0       # AR_STEPS LLM steps + RF samples, AR_STEPS - 1 sem-decoder steps
        t["sem_steps"] = sem_t[0]
        # the LLM part of one AR step alone (to scale 2 layers -> 28): one more cached step on B = 3 rows
        for c in caches:
            c["k"], c["v"] = c["k"].repeat(CFG_ROWS, 1, 1, 1), c["v"].repeat(CFG_ROWS, 1, 1, 1)
        am = torch.ones((CFG_ROWS, caches[0]["k"].shape[2] + 1), dtype=torch.long)
        c0 = time.perf_counter()
        L.model_forward(sd, cfg, start.repeat(CFG_ROWS, 1, 1), am, (am.cumsum(-1) - 1)[:, -1:], caches)
        t["llm_step_2layers"] = time.perf_counter() - c0
        c0 = time.perf_counter()
        # the pixel decoder sees the round's 64 generated tokens; the sample feeds it the 64 ENCODED features instead
        recon = O.pixel_decoder_forward(self.tok_sd, enc["x_norm_patchtokens"], self.tok_cfg["semantic_decoder"],
                                        self.tok_cfg["pixel_decoder"])
        t["pix"] = time.perf_counter() - c0
        if keep:
            t["_keep"] = dict(hidden_last=h[:, -1], lat0=lats[0] if lats else None, feats=enc["x_norm_patchtokens"],
                              recon=recon, noises=noises)
        return t

    def extrapolate(self, t: dict) -> dict:
        """Full-round seconds from the sampled stages: LLM parts x 28 / 2 layers, AR steps x 65 / sampled."""
        f = 28.0 / self.LAYERS
        n_ar = N_GEN + 1
        sem_step = t["sem_steps"] / max(self.AR_STEPS - 1, 1)
        rf_and_heads = (t["ar_steps"] - t["sem_steps"]) / self.AR_STEPS - t["llm_step_2layers"]
        per_tok = f * t["llm_step_2layers"] + max(rf_and_heads, 0.0) + sem_step
        full = t["enc"] + f * t["prefill_2layers"] + n_ar * per_tok + t["pix"]
        return {"full_round_s": full, "per_ar_step_s": per_tok, "llm_step_28layers_s": f * t["llm_step_2layers"],
                "rf_vis_head_s": max(rf_and_heads, 0.0), "sem_step_s": sem_step, "enc_s": t["enc"],
                "prefill_28layers_s": f * t["prefill_2layers"], "pix_s": t["pix"]}

    SAMPLE = ("one 256x256 edit round with the LLM depth-reduced to 2 of 28 layers (true widths, 64 experts) and 2 of the "
              "65 AR steps; fp32 torch CPU restatement; stages timed separately and extrapolated: LLM x14, AR steps x32.5")


def run_reference_arm(args, world, rank):
    """`--impl reference`: the reference's own CPU implementation of the path.  The reference is pure Python/PyTorch
    and cannot travel to the GPU box, so this is the oracle port (kind "port"); each step is one bounded sample."""
    if rank != 0:
        return
    ref = CpuReference()
    for _ in range(min(args.warmup, 1)):
        ref.sample_step()
    ts, walls = [], []
    for _ in range(args.steps):
        w0 = time.perf_counter()
        ts.append(ref.sample_step())
        walls.append(time.perf_counter() - w0)
    med = {k: statistics.median(t[k] for t in ts) for k in ts[0]}
    ex = ref.extrapolate(med)
    value = TOKENS_PER_STEP / ex["full_round_s"]
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * statistics.mean(walls),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "tokens_per_step": TOKENS_PER_STEP, "sample": ref.SAMPLE,
                       "extrapolated": True, "extrapolated_full_round_ms": 1e3 * ex["full_round_s"],
                       "stages_s": {k: round(v, 4) for k, v in ex.items()},
                       "note": "ms_per_step is the measured wall time of one bounded sample step; value = 128 tokens / "
                               "the full-round time extrapolated from that sample's stage timings"},
            "generated_tokens_per_s": N_GEN / ex["full_round_s"],
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": ref.cores, "kind": "port", "sample": ref.SAMPLE,
                             "extrapolated": True},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------------
def build_full_model(dev, rank: int, world: int):
    from ming_univision_b200 import synthetic
    from ming_univision_b200.mingtok import MingTokConfig
    from ming_univision_b200.modeling_bailing_moe import BailingMoeConfig
    from ming_univision_b200.modeling_bailingmm import MingUniVisionForConditionalGeneration

    llm_cfg = dict(synthetic.LLM_CONFIG, num_image_tokens_for_gen=N_GEN)
    m = MingUniVisionForConditionalGeneration.on_device(BailingMoeConfig(**llm_cfg), MingTokConfig(**synthetic.MINGTOK_CONFIG),
                                                        synthetic.VISHEAD_CONFIG, dev, ep_rank=rank, ep_size=world)
    synthetic.init_on_device(m, seed=0)
    if world > 1:
        m.model.model.set_expert_parallel(None, mode="dispatch", t_max=128)
    return m, llm_cfg


def measure_mingtok_stage(vision, dev, steps: int = 5, warmup: int = 3) -> dict:
    """BASELINE configs[1] (last round's headline) as a secondary stage: MingTok enc+dec, batch 64, 256x256, bf16."""
    from ming_univision_b200 import ops, synthetic

    imgs = [synthetic.synthetic_images(MINGTOK_BATCH, SIZE, seed=1234 + i).to(dev).to(torch.bfloat16) for i in range(3)]
    for i in range(warmup):
        vision.forward_enc_dec(imgs[i % 3])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        vision.forward_enc_dec(imgs[i % 3])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    ops.REPLAY = []
    vision.forward_enc_dec(imgs[0])
    torch.cuda.synchronize()
    replay, ops.REPLAY = ops.REPLAY, None
    flops = sum(f for _, f, _ in replay)
    for fn, _, _ in replay:
        fn()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(3):
        for fn, _, _ in replay:
            fn()
    e1.record()
    torch.cuda.synchronize()
    gemm_ms = e0.elapsed_time(e1) / 3
    peaks = _peaks()
    tokens = MINGTOK_BATCH * N_ENC
    return {"workload": "MingTok ViT enc+dec batch=64 256x256 bf16 (BASELINE configs[1])", "tokens_per_s": tokens / (ms / 1e3),
            "ms_per_step": ms, "steps": steps, "step_tflops_algorithmic": 213.0e9 * MINGTOK_BATCH / (ms / 1e3) / 1e12,
            "roofline": {"bound": "tensor", "kernel": "mb::gemm_bf16_kernel (tcgen05.mma + TMA)",
                         "achieved": flops / (gemm_ms / 1e3) / 1e12, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                         "frac": flops / (gemm_ms / 1e3) / 1e12 / peaks["bf16_tflops"], "gemm_ms_per_step": gemm_ms,
                         "gemm_share_of_step": gemm_ms / ms, "gemm_launches_per_step": len(replay),
                         "timing": "replay of the step's GEMM launches, CUDA events"}}


def run_ours(args, world, rank, local):
    from ming_univision_b200 import _lib, ops

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    _lib.require_device()
    t_build = time.perf_counter()
    model, llm_cfg = build_full_model(dev, rank, world)
    llm = model.model
    torch.cuda.synchronize()
    t_build = time.perf_counter() - t_build
    mem_gb = torch.cuda.memory_allocated(dev) / 1e9

    # several distinct rounds are cycled (and every rank has its own): inputs differ step to step; the 38 GB of weights
    # a round streams 65 times exceed the 126 MB L2 by far, so nothing a step reads is left over from the previous one
    n_bufs = 3
    NI = args.images
    if NI < 1 or NI * CFG_ROWS > 6:
        raise SystemExit("--images: 1 or 2 rounds per step (images x 3 CFG rows <= 6 rows of the weight-streaming kernels)")
    rounds = [[round_inputs(llm_cfg, 1000 * rank + 10 * i + j) for j in range(NI)] for i in range(n_bufs)]
    ids_host = [torch.cat([r[0] for r in rs]).pin_memory() for rs in rounds]                 # [NI, 104]
    img_host = [torch.stack([r[3] for r in rs]).pin_memory() for rs in rounds]               # [NI, 256, 256, 3] u8
    um = rounds[0][0][1].to(dev).expand(NI, -1).contiguous()
    tm = rounds[0][0][2].to(dev).expand(NI, -1).contiguous()
    ids_dev = [t.to(dev) for t in ids_host]
    px_dev = [ops.image_preprocess(t.to(dev), SIZE, SIZE, out_dtype=torch.bfloat16) for t in img_host]

    def step_resident(i):
        img, _ = model.generate_image_from_prompt(ids_dev[i % n_bufs], pixel_values=px_dev[i % n_bufs],
                                                  uncond_attention_mask=um, text_uncond_attention_mask=tm)
        return img

    host_out = torch.empty((NI, SIZE, SIZE, 3), dtype=torch.uint8).pin_memory()

    def step_e2e(i):
        ids = ids_host[i % n_bufs].to(dev, non_blocking=True)
        u8 = img_host[i % n_bufs].to(dev, non_blocking=True)
        px = ops.image_preprocess(u8, SIZE, SIZE, out_dtype=torch.bfloat16)   # Resize(256) + CenterCrop + Normalize
        img, _ = model.generate_image_from_prompt(ids, pixel_values=px, uncond_attention_mask=um,
                                                  text_uncond_attention_mask=tm)
        host_out.copy_(ops.image_postprocess(img), non_blocking=True)          # tensor_to_pil's u8 conversion
        torch.cuda.current_stream().synchronize()                              # the caller holds the finished image
        return host_out

    for i in range(args.warmup):
        step_resident(i)
    step_e2e(0)
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
    for i in range(args.steps):
        step_e2e(i)
    e1.record()
    torch.cuda.synchronize()
    _barrier(world)
    ms_e2e = _max_over_ranks(e0.elapsed_time(e1), world, dev)
    clocks = sampler.stop() if sampler else None
    if world > 1:
        llm.model.ep_peer.check()

    # ---- per-stage split of one round (all ranks: the MoE layers exchange rows, so every rank makes the same calls)
    def timed(fn):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        r = fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b), r

    ms_round, _ = timed(lambda: step_resident(0))
    ms_enc, _ = timed(lambda: model.extract_image_feature(px_dev[0]))
    feats64 = model.vision(px_dev[0])["x_norm_patchtokens"]                   # [NI, 64, 1024]
    ms_pix, _ = timed(lambda: model.vision.forward_pixel_decoder(feats64))
    z = torch.randn((CFG_ROWS * NI, 3072), device=dev).to(torch.bfloat16)
    llm.diffloss.sample(z, 1.0, 3.0, 1.1, groups=NI)
    ms_rf, _ = timed(lambda: [llm.diffloss.sample(z, 1.0, 3.0, 1.1, groups=NI) for _ in range(8)])
    ms_rf /= 8

    # ---- the same round ONE request at a time (the reference's batch-1 semantics, modeling_bailing_moe.py:1865) — all
    # ranks run it (the MoE layers exchange rows in lockstep)
    single = None
    if NI > 1:
        def step_single(i):
            return model.generate_image_from_prompt(ids_dev[i % n_bufs][0:1], pixel_values=px_dev[i % n_bufs][0:1],
                                                    uncond_attention_mask=um[0:1], text_uncond_attention_mask=tm[0:1])[0]

        step_single(0)
        ms_single, _ = timed(lambda: [step_single(i) for i in range(3)])
        ms_single /= 3
        single = {"ms_per_round": ms_single, "tokens_per_s": TOKENS_PER_STEP / (ms_single / 1e3),
                  "generated_tokens_per_s": N_GEN / (ms_single / 1e3), "steps": 3,
                  "note": "one edit round per step and GPU (3 rows per pass over the weights instead of 6)"}
        step_resident(0)

    # ---- N > 1: consistency check of the peer-memory exchange on the hardware that has the peers (tests/test_ep_gpu.py
    # needs >= 2 GPUs too and is skipped on a 1-GPU box).  Every rank generates the SAME request (same prompt, image and RF
    # noise): its rows travel to the same expert owners and the partial sums are added in rank order, so every rank must
    # end with the same image, bit for bit.
    ep_check = None
    if world > 1:
        import torch.distributed as dist

        probe = round_inputs(llm_cfg, 424242)
        gen = torch.Generator().manual_seed(7)
        probe_noises = [torch.randn((1, 32), generator=gen) for _ in range(N_GEN + 1)]
        probe_px = ops.image_preprocess(probe[3].unsqueeze(0).to(dev), SIZE, SIZE, out_dtype=torch.bfloat16)
        probe_img, _ = model.generate_image_from_prompt(probe[0].to(dev), pixel_values=probe_px,
                                                        uncond_attention_mask=probe[1].to(dev),
                                                        text_uncond_attention_mask=probe[2].to(dev), noises=probe_noises)
        u8 = ops.image_postprocess(probe_img).to(torch.int64).flatten()
        weights = torch.arange(u8.numel(), device=dev, dtype=torch.int64) % 8191 + 1
        digest = torch.stack((u8.sum(), (u8 * weights).sum()))
        digests = [torch.zeros_like(digest) for _ in range(world)]
        dist.all_gather(digests, digest)
        ep_check = {"what": "the same request generated on every rank (expert-parallel dispatch / combine): image digests",
                    "identical_on_all_ranks": all(bool(torch.equal(d, digests[0])) for d in digests), "ranks": world}
        step_resident(0)

    # ---- roofline of the dominant kernel, the persistent RF sampler (one launch per generated token: 16 Euler steps x 12
    # residual blocks): CUDA events around its launches on the eager path (launching stream), algorithmic bytes = the
    # bf16 weights it must stream = steps x depth x 3 H W x 2
    from ming_univision_b200 import diff_loss_rf_swiglu as rfmod

    llm.diffloss.use_cuda_graph = False
    llm.diffloss.sample(z, 1.0, 3.0, 1.1, groups=NI)
    rfmod.FUSED_PROFILE = []
    for _ in range(5):
        llm.diffloss.sample(z, 1.0, 3.0, 1.1, groups=NI)
    torch.cuda.synchronize()
    rf_prof, rfmod.FUSED_PROFILE = rfmod.FUSED_PROFILE, None
    llm.diffloss.use_cuda_graph = True
    rf_bytes = rf_prof[0][0] if rf_prof else 0
    rf_kernel_ms = statistics.median(a.elapsed_time(b) for _, a, b in rf_prof) if rf_prof else float("nan")

    # ---- second: every launch of the per-layer weight-streaming kernel (gemv: LLM dense layers, semantic decoder, heads)
    # of TWO token steps is recorded on the eager path and the list re-issued back to back between two CUDA events
    cfg_obj = llm.config
    saved = (cfg_obj.num_image_tokens_for_gen, llm.use_cuda_graph)
    cfg_obj.num_image_tokens_for_gen, llm.use_cuda_graph = 1, False
    ops.GEMV_REPLAY = []
    step_resident(0)
    torch.cuda.synchronize()
    replay, ops.GEMV_REPLAY = ops.GEMV_REPLAY, None
    cfg_obj.num_image_tokens_for_gen, llm.use_cuda_graph = saved
    n_token_steps = 2
    for fn, _, _ in replay:
        fn()
    ms_gemv, _ = timed(lambda: [fn() for _ in range(3) for fn, _, _ in replay])
    ms_gemv /= 3
    gemv_bytes = sum(b for _, b, _ in replay)
    step_resident(0)  # the graph path still works after the eager detour

    if rank != 0:
        return
    peaks = _peaks()
    traffic, traffic_src = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "rf_fused_traffic.json")) as f:
            tj = json.load(f)
        traffic, traffic_src = tj["dram_bytes_per_launch"], tj["source"]
    except (OSError, KeyError, ValueError):
        pass
    ms_step = ms_res / args.steps
    ms_ar_step = (ms_round - ms_enc - ms_pix) / (N_GEN + 1)     # includes the prefill's share (one prefill per round)
    achieved = rf_bytes / (rf_kernel_ms / 1e3) / 1e9
    gemv_ms_per_token = ms_gemv / n_token_steps
    roofline = {"bound": "hbm", "kernel": "mb::rf_sample_fused_kernel (persistent weight streaming: the RF head's 16 Euler "
                                          "steps x 12 residual blocks, cp.async.bulk ring + mma.sync)",
                "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
                "peak_source": peaks["source"], "traffic": traffic, "traffic_unit": "bytes per launch (dram read + write)",
                "traffic_source": traffic_src, "algorithmic_bytes_per_launch": rf_bytes, "launch_ms": rf_kernel_ms,
                "launches_per_token_step": 1, "rows_per_launch": CFG_ROWS * NI,
                "share_of_token_step": rf_kernel_ms / ms_ar_step, "token_step_ms": ms_ar_step,
                "timing": "CUDA events around the kernel's launches (eager path, launching stream), median of 5",
                "token_step_gbs_survey_bytes": 46.0e9 * NI / (ms_ar_step / 1e3) / 1e9,
                "second_kernel": {"kernel": "mb::gemv_bf16_kernel (per-layer weight streaming: LLM dense layers, semantic "
                                            "decoder, heads)",
                                  "achieved": gemv_bytes / (ms_gemv / 1e3) / 1e9, "unit": "GB/s",
                                  "frac": gemv_bytes / (ms_gemv / 1e3) / 1e9 / peaks["hbm_gbs"],
                                  "launches_per_token_step": len(replay) // n_token_steps,
                                  "bytes_per_token_step": gemv_bytes / n_token_steps,
                                  "ms_per_token_step": gemv_ms_per_token,
                                  "share_of_token_step": gemv_ms_per_token / ms_ar_step,
                                  "timing": "replay of two token steps' launches back to back, CUDA events"},
                "note": "algorithmic bytes = the bf16 weights a launch must stream; the token step also streams the routed "
                        "experts (mb::moe_expert_kernel, <= 18 x 17.3 MB x 28 layers per 3 rows) and the KV caches; "
                        "token_step_gbs_survey_bytes uses SURVEY.md §8d's 46 GB per token and request (RF without the "
                        "adaLN hoist) x the requests generated together"}
    stages = {"single_request": single, "round_ms": ms_round, "enc_ms": ms_enc, "pixel_decoder_ms": ms_pix, "rf_sample_ms": ms_rf,
              "ar_step_ms_incl_prefill_share": ms_ar_step,
              "mingtok_enc_dec": measure_mingtok_stage(model.vision, dev)}

    # ---- CPU baseline (oracle port, bounded sample, extrapolated) + parity of that sample against the CUDA path
    del feats64
    cpu, parity = None, None
    if not args.no_cpu and world == 1:  # rank 0 at N = 1 only (the contract); N > 1 lines carry null
        ref = CpuReference()
        t = ref.sample_step(keep=True)
        ex = ref.extrapolate({k: v for k, v in t.items() if not k.startswith("_")})
        cpu = {"value": TOKENS_PER_STEP / ex["full_round_s"], "unit": UNIT, "cores": ref.cores, "kind": "port",
               "sample": ref.SAMPLE, "extrapolated": True, "stages_s": {k: round(v, 4) for k, v in ex.items()}}
        parity = parity_vs_oracle(ref, t["_keep"], dev)

    tokens_all = TOKENS_PER_STEP * world * NI
    line = {"metric": METRIC, "value": tokens_all * args.steps / (ms_res / 1e3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "generated_tokens_per_s": N_GEN * world * NI * args.steps / (ms_res / 1e3),
            "config": {"workload": WORKLOAD, "tokens_per_step": tokens_all, "encoded_tokens_per_round": N_ENC,
                       "generated_tokens_per_round": N_GEN, "cfg_rows": CFG_ROWS, "rounds_per_step": world * NI,
                       "rounds_generated_together_per_gpu": NI,
                       "parallelism": (f"dp{world} x ep{world}: every GPU generates its own {NI} edit round(s), routed experts sharded "
                                       f"{64 // world} per GPU, dispatch / combine through NVLink peer memory in-kernel "
                                       "(no NCCL on the data path)") if world > 1 else "1 GPU, all 64 experts resident",
                       "l2_policy": "3 distinct rounds cycled; the 38 GB of weights streamed per AR step exceed the 126 MB L2",
                       "model_build_s": round(t_build, 1), "weights_gb_per_gpu": round(mem_gb, 1)},
            "e2e": {"value": tokens_all * args.steps / (ms_e2e / 1e3), "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": NI * (SIZE * SIZE * 3 + PROMPT_LEN * 8),
                    "d2h_bytes_per_step": NI * SIZE * SIZE * 3},
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "stages": stages,
            "cpu_baseline": cpu, "parity": parity, "ep_check": ep_check}
    print(json.dumps(line), flush=True)


def parity_vs_oracle(ref: "CpuReference", keep: dict, dev) -> dict:
    """The CUDA path on the SAME 2-layer true-width weights / inputs as the CPU sample: MingTok features, prefill hidden
    state, the first generated latent (one full RF sample behind a real prefill) and the reconstruction."""
    from ming_univision_b200.mingtok import MingTokConfig
    from ming_univision_b200.modeling_bailing_moe import BailingMoeConfig
    from ming_univision_b200.modeling_bailingmm import MingUniVisionForConditionalGeneration

    m = MingUniVisionForConditionalGeneration.on_device(BailingMoeConfig(**ref.cfg), MingTokConfig(**ref.tok_cfg), ref.vh, dev)
    sd = {(k if k.startswith("linear_proj.") else "model." + k): v for k, v in ref.sd.items()}
    sd.update({"vision." + k: v for k, v in ref.tok_sd.items()})
    m.load_state_dict(sd, strict=True)
    from ming_univision_b200 import ops

    px = ops.image_preprocess(ref.img_u8.to(dev), SIZE, SIZE, out_dtype=torch.bfloat16)
    feats = m.vision(px)["x_norm_patchtokens"]
    lats = []
    orig = m.vision.forward_feature_decoder

    def spy(latent, past_key_values=None):
        lats.append(latent.float().cpu())
        return orig(latent, past_key_values=past_key_values)

    m.vision.forward_feature_decoder = spy
    m.model.config.num_image_tokens_for_gen = 1
    m.generate_image_from_prompt(ref.ids.to(dev), pixel_values=px, uncond_attention_mask=ref.um.to(dev),
                                 text_uncond_attention_mask=ref.tm.to(dev), noises=keep["noises"])
    m.vision.forward_feature_decoder = orig
    recon = m.vision.forward_pixel_decoder(feats, out_dtype=torch.float32).cpu()

    def rel(a, b):
        a, b = a.double().cpu(), b.double().cpu()
        return float((a - b).norm() / b.norm())

    img = (ref.img_u8.float() / 255.0 - 0.5) / 0.5
    img = img.permute(2, 0, 1).unsqueeze(0)
    mse_o, mse_r = float(((recon - img) ** 2).mean()), float(((keep["recon"] - img) ** 2).mean())
    return {"what": "CUDA path vs the fp32 CPU oracle on the sampled round (2-layer true-width LLM, full RF head, full MingTok)",
            "rel_l2_mingtok_features": rel(feats, keep["feats"]),
            "rel_l2_first_generated_latent": rel(lats[0][0:1], keep["lat0"][0:1]) if keep["lat0"] is not None else None,
            "rel_l2_recon": rel(recon, keep["recon"]),
            "dpsnr_vs_input_db": abs(10 * math.log10(4.0 / mse_o) - 10 * math.log10(4.0 / mse_r))}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline / parity leg (development runs)")
    ap.add_argument("--images", type=int, default=int(os.environ.get("MB_BENCH_IMAGES", "2")),
                    help="edit rounds generated TOGETHER per GPU (batched serving; images x 3 CFG rows <= 6); the "
                         "single-request number (the reference's batch 1) is reported next to it")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":  # CPU arm: rank 0 alone works, no process group needed
        run_reference_arm(args, int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")))
        return
    world, rank, local = _dist_setup()
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
