#!/bin/bash
# Round-2 ncu evidence for profiles/ (run under gpurun, one GPU):
#  (1) launch list (per-launch device time) of the first kernels of bench.py's timed region — MB_NCU_RANGE=1 brackets it
#      with cudaProfilerStart/Stop; kernels inside CUDA-graph replays are listed individually;
#  (2) --set full of the persistent RF sampler kernel at the bench's row count (6 rows = 2 requests x 3 CFG rows);
#  (3) --set full of the AR step's other weight-streaming kernels on the eager path: gemv (LLM qkv / shared experts),
#      moe_expert_kernel (gate-up, down), GQA decode attention.
set -u
mkdir -p gpurun_out
TAG=${1:-r02}
MB_NCU_RANGE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv -c 4000 \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_bench_${TAG}.log 2>&1
RF_ROWS=6 timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:rf_sample_fused \
    -o gpurun_out/prof_rf_fused_${TAG} -f python tools/profile_rf.py > gpurun_out/ncu_rf_${TAG}.log 2>&1
# one token step behind the prefill: skip the prefill / encoder launches by kernel-name filters
AR_LAYERS=4 timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k regex:"moe_expert_kernel|attn_decode_gqa128" -s 8 -c 6 -o gpurun_out/prof_ar_moe_${TAG} -f \
    python tools/profile_ar_step.py > gpurun_out/ncu_ar_moe_${TAG}.log 2>&1
AR_LAYERS=4 timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k regex:gemv_bf16_kernel -s 10 -c 8 -o gpurun_out/prof_ar_gemv_${TAG} -f \
    python tools/profile_ar_step.py > gpurun_out/ncu_ar_gemv_${TAG}.log 2>&1
ls -la gpurun_out/ | grep ${TAG}
