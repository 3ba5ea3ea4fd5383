#!/bin/bash
# ncu evidence for profiles/: (1) launch list with per-launch device time of ~2 bench steps, (2) --set full capture
# of the four pixel-decoder GEMMs (qkv, proj+residual, fc1+GELU, fc2+residual) of one step.  Run under gpurun.
set -u
mkdir -p gpurun_out
TAG=${1:-r01}
# one forward_enc_dec = 434 launches of ours; bench warm-up = 3 x (resident + e2e) passes = 2604 launches
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2604 -c 900 --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_bench_${TAG}.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 147 -c 4 \
    -o gpurun_out/prof_gemm_${TAG} -f python bench.py --steps 1 --warmup 3 > gpurun_out/ncu_full_${TAG}.log 2>&1
ls -la gpurun_out/ | tail -8
