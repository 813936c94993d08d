#!/bin/bash
# ncu captures of the round-2 default path (run under gpurun, ONE GPU).   tools/profile_r02.sh <tag>
TAG=${1:-r02}
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
NCU="ncu --clock-control none"
# 1) launch list with per-launch device time of ONE eagerly enqueued timed step (shares, not absolutes: ncu serialises, cold caches)
PN_CUDA_PROFILER=1 $NCU --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-stock-torch > gpurun_out/${TAG}_launches_bench.log 2>&1
# 2) --set full: the fused loss (forward + backward launch of the third iteration)
$NCU --set full --import-source on -k regex:"loss_group_kernel" -s 4 -c 2 -f -o gpurun_out/${TAG}_loss python tools/loss_only.py > gpurun_out/${TAG}_loss_ncu.log 2>&1
# 3) --set full: folded pack1 forward (256 -> 64, 7x7 at 96x320), its data gradient and weight gradient
$NCU --set full --import-source on -k regex:"conv_igemm_kernel" -s 2 -c 1 -f -o gpurun_out/${TAG}_conv_pack1f python tools/conv_only.py 4 96 320 256 64 7 fwd > gpurun_out/${TAG}_conv_ncu.log 2>&1
$NCU --set full --import-source on -k regex:"conv_igemm_kernel" -s 2 -c 1 -f -o gpurun_out/${TAG}_dgrad_pack1f python tools/conv_only.py 4 96 320 256 64 7 dgrad >> gpurun_out/${TAG}_conv_ncu.log 2>&1
$NCU --set full --import-source on -k regex:"conv_wgrad_kernel" -s 2 -c 1 -f -o gpurun_out/${TAG}_wgrad_pack1f python tools/conv_only.py 4 96 320 256 64 7 wgrad >> gpurun_out/${TAG}_conv_ncu.log 2>&1
# 4) a small-map layer (12x40, 512 -> 512, 3x3: 10 calls per step at 90 TFLOP/s)
$NCU --set full --import-source on -k regex:"conv_igemm_kernel" -s 2 -c 1 -f -o gpurun_out/${TAG}_conv_12x40 python tools/conv_only.py 4 12 40 512 512 3 fwd >> gpurun_out/${TAG}_conv_ncu.log 2>&1
$NCU --set full --import-source on -k regex:"conv_wgrad_kernel" -s 2 -c 1 -f -o gpurun_out/${TAG}_wgrad_12x40 python tools/conv_only.py 4 12 40 512 512 3 wgrad >> gpurun_out/${TAG}_conv_ncu.log 2>&1
ls -la gpurun_out | tail -12
