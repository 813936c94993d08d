#!/bin/bash
# First GPU call of round 2: everything written after the round-1 GPU budget ran out, in one gpurun.
#   gpurun --timeout 3000 -- 'bash tools/round2_first_call.sh'      (about 35-45 minutes of box time)
# 1) default -m gpu tier (the measured path must still be green)
# 2) experimental tiers (PN_EXPERIMENTAL=1): folded pack block + CUDA graph, grouped-scale loss program, re-compositions /
#    flat staging / GroupNorm tree / tiled unpack (tests/test_recompose_gpu.py)
# 3) kernel timings: loss tile vs grouped, head convolution default vs flat staging
# 4) bench lines: default, --staged-small, --pack-fold, --staged-all, --graph, --graph --staged-all, then one per variant
# 5) CUPTI step breakdowns, conv tile-height sweep, the 384x1280 B=2 shape
# Read gpurun_out/r02a_* afterwards; DESIGN.md 7.4 says which result flips which default.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
O=gpurun_out/r02a
timeout 1200 python -m pytest tests -m gpu -x -q                                   > ${O}_tests_default.log 2>&1; echo "default tier: $?"
PN_EXPERIMENTAL=1 timeout 900 python -m pytest tests/test_folded_gpu.py tests/test_graph_gpu.py -m gpu -q  > ${O}_tests_folded.log 2>&1;  echo "folded tier: $?"
PN_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_loss_gpu.py -m gpu -q -k grouped > ${O}_tests_loss_grouped.log 2>&1;  echo "grouped loss tier: $?"
PN_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_recompose_gpu.py -m gpu -q -s > ${O}_tests_recompose.log 2>&1;  echo "recompose tier: $?"; tail -3 ${O}_tests_recompose.log
tail -5 ${O}_tests_default.log ${O}_tests_folded.log ${O}_tests_loss_grouped.log
timeout 300 python tools/loss_only.py > ${O}_loss_only_tile.txt 2>&1; PN_LOSS_GROUPED=1 timeout 300 python tools/loss_only.py > ${O}_loss_only_grouped.txt 2>&1
tail -3 ${O}_loss_only_tile.txt ${O}_loss_only_grouped.txt
timeout 300 python tools/head_bench.py > ${O}_head_bench.txt 2>&1; tail -8 ${O}_head_bench.txt
# the combined runs first (they decide the defaults), then one line per variant
for flags in "" "--staged-small" "--pack-fold" "--staged-all" "--graph" "--graph --staged-all" "--loss-grouped" "--im2col-first" "--stage-flat" "--gn-tree" "--unpack-tiled" "--pack-tiled"; do
  tag=$(echo "default $flags" | tr -d ' -' )
  timeout 600 python bench.py --no-cpu-baseline --no-staged-probe $flags > ${O}_bench_${tag}.log 2> ${O}_bench_${tag}.err
  echo "bench [$flags]: $? $(cut -c1-400 ${O}_bench_${tag}.log)"
done
timeout 300 python tools/step_profile.py             > ${O}_step_breakdown.txt 2>&1
timeout 300 python tools/step_profile.py --pack-fold > ${O}_step_breakdown_fold.txt 2>&1
timeout 300 python tools/step_profile.py --staged-small > ${O}_step_breakdown_staged_small.txt 2>&1
timeout 300 python tools/step_profile.py --staged-all > ${O}_step_breakdown_staged_all.txt 2>&1
head -30 ${O}_step_breakdown_fold.txt
timeout 600 python tools/conv_sweep.py > ${O}_conv_sweep.txt 2>&1; tail -5 ${O}_conv_sweep.txt
# BASELINE configs[2] shape on one GPU (B=2, 384x1280): never run on the B200 in round 1
timeout 600 python bench.py --no-cpu-baseline --height 384 --width 1280 --batch 2 > ${O}_bench_384x1280.log 2> ${O}_bench_384x1280.err
echo "bench [384x1280 B=2]: $? $(cut -c1-300 ${O}_bench_384x1280.log)"
