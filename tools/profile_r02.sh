#!/bin/bash
# ncu captures of the staged round-2 path (run under gpurun, ONE GPU, after tools/round2_first_call.sh is green).
#   tools/profile_r02.sh <tag>
TAG=${1:-r02}
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
# launch list of one timed step with the folded pack layers (shares, not absolutes: ncu serialises and runs cold)
PN_CUDA_PROFILER=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/${TAG}_fold_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --pack-fold > gpurun_out/${TAG}_fold_launches_bench.log 2>&1
# --set full of the folded pack1 block: the 7x7 convolution (fprop, dgrad), its weight gradient, the fold and frame kernels
ncu --set full --clock-control none --import-source on -k regex:"conv_igemm_kernel|conv_wgrad_kernel|fold_fwd_kernel|fold_bwd_kernel|frame_" \
    -s 50 -c 25 -f -o gpurun_out/${TAG}_folded_pack1 python tools/folded_only.py > gpurun_out/${TAG}_folded_ncu.log 2>&1
ls -la gpurun_out | tail -8
# --set full of the grouped-scale loss program (forward + backward launches of the third iteration)
PN_LOSS_GROUPED=1 ncu --set full --clock-control none --import-source on -k regex:"loss_group_kernel" \
    -s 4 -c 2 -f -o gpurun_out/${TAG}_loss_grouped python tools/loss_only.py > gpurun_out/${TAG}_loss_grouped_ncu.log 2>&1
ls -la gpurun_out | tail -4
