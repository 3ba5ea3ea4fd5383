#!/bin/bash
# NVLink evidence for the expert-parallel exchange kernels: a 2-rank job (one process per GPU) in which ONLY rank 0 runs under
# ncu (kernel replay of the push kernels is idempotent: same rows, same epoch flag), rank 1 runs plainly and waits for it.
# Run under `gpurun --gpus 2`.
set -u
mkdir -p gpurun_out
export MASTER_ADDR=127.0.0.1 MASTER_PORT=29541 WORLD_SIZE=2 EP_LAYERS=40 MB_EP_TIMEOUT_MS=120000
RANK=1 LOCAL_RANK=1 python tools/ep_nvlink_traffic.py > gpurun_out/ep_push_rank1.log 2>&1 &
P1=$!
RANK=0 LOCAL_RANK=0 timeout 500 ncu --set full --clock-control none --import-source on \
    -k regex:"ep_combine_push_kernel|ep_dispatch_push_kernel|ep_reduce_finalize_kernel" -s 6 -c 3 \
    -o gpurun_out/prof_ep_push_r02 -f python tools/ep_nvlink_traffic.py > gpurun_out/ep_push_rank0.log 2>&1
wait $P1
tail -3 gpurun_out/ep_push_rank0.log
ncu -i gpurun_out/prof_ep_push_r02.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
r=list(csv.reader(sys.stdin)); h=r[0]
for i,n in enumerate(h):
    if 'nvl' in n.lower() or n in ('Kernel Name','gpu__time_duration.sum','lts__t_sectors_srcunit_tex_aperture_peer.sum','lts__t_sectors_aperture_peer.sum'):
        print(n, [row[i] for row in r[1:5]])
" | head -40
