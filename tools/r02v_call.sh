#!/bin/bash
# frame GEMM pipeline, unpack stencil forward stores, stencil backward prefetch, GN cluster fwd unroll 8: tests, A/B benches, CUPTI breakdown
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
O=gpurun_out/r02v
SECONDS=0
timeout 900 python -m pytest tests/test_recompose_gpu.py tests/test_layers_gpu.py tests/test_packnet_gpu.py tests/test_folded_gpu.py tests/test_loss_gpu.py tests/test_step_gpu.py -m gpu -q -s > ${O}_tests.log 2>&1; echo "gpu tests: rc $? in ${SECONDS}s"; grep -E "passed|failed|FAILED|Error" ${O}_tests.log | cut -c1-300 | head; grep POSE_REL ${O}_tests.log | sort | uniq | awk '{print $2,$3,$4}' | sort -k3 -g | tail -8
grep -E "disp[1-4] .*max-rel|worst parameter" ${O}_tests.log | tail -12
for v in "" "PN_STENCIL_PREFETCH=0" "PN_GN_CLUSTER_MAX_MB=34"; do
  env $v timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-stock-torch > ${O}_bench.log 2> ${O}_bench.err
  echo "[$v] $(grep -o '"ms_per_step": [0-9.]*' ${O}_bench.log | head -1)"
done
timeout 300 python tools/step_profile.py --steps 2 --top 60 > ${O}_step_profile.log 2>&1; grep -E "device time|frame|stencil|gn_|head_" ${O}_step_profile.log | cut -c1-150
timeout 300 python tools/gn_bench.py 0 1:34 > ${O}_gn_bench.log 2>&1; grep -v Warn ${O}_gn_bench.log | grep x | head -12
