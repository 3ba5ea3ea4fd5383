#!/bin/bash
# ncu evidence for profiles/: (1) launch list with per-launch device time of the timed region of bench.py (2 steps;
# MB_NCU_RANGE=1 brackets it with cudaProfilerStart/Stop), (2) --set full capture of the four pixel-decoder GEMMs
# (qkv, proj+residual, fc1+GELU, fc2+residual) of the first pass.  Run under gpurun.
set -u
mkdir -p gpurun_out
TAG=${1:-r01}
MB_NCU_RANGE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_bench_${TAG}.log 2>&1
# GEMM launches of one pass: 1 patch + 48 enc + 1 out + 96 sem + 1 sem_to_pix = 147 before the first pixel block
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 147 -c 4 \
    -o gpurun_out/prof_gemm_${TAG} -f python bench.py --steps 1 --warmup 3 > gpurun_out/ncu_full_${TAG}.log 2>&1
ls -la gpurun_out/ | tail -8
# tcgen05 attention kernel: one pixel-decoder launch (12 encoder + 24 semantic-decoder launches come first in a pass)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_tc -s 40 -c 1 \
    -o gpurun_out/prof_attn_${TAG} -f python bench.py --steps 1 --warmup 3 > gpurun_out/ncu_attn_${TAG}.log 2>&1
ls -la gpurun_out/ | tail -4
# image pre-processing kernels: --set full of the horizontal and vertical passes on the 64-image batch
timeout 600 ncu --set full --clock-control none --import-source on -k regex:resample_ -c 6 \
    -o gpurun_out/prof_preprocess_${TAG} -f python tools/bench_preprocess.py --iters 1 > gpurun_out/ncu_preprocess_${TAG}.log 2>&1
ls -la gpurun_out/ | tail -3
