#!/bin/bash
# ncu captures for profiles/ (run under gpurun, ONE GPU).  Usage: tools/profile.sh <tag>
# 1) launch list with per-launch device time of the TIMED step of one bench run (cold-cache, serialised: compare SHARES)
# 2) --set full capture of the two hot kernels
TAG=${1:-r01}
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
PN_CUDA_PROFILER=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:loss_tile_kernel -s 2 -c 2 -f -o gpurun_out/${TAG}_loss \
    python tools/loss_only.py > gpurun_out/${TAG}_loss_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_igemm_kernel -s 1 -c 1 -f -o gpurun_out/${TAG}_conv \
    python tools/conv_only.py > gpurun_out/${TAG}_conv_ncu.log 2>&1
ls -la gpurun_out | tail -12
