#!/bin/bash
# cp.async double buffering of the stencil backward kernels + in-place gradient windows: tests, A/B, breakdown
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
O=gpurun_out/r02w
SECONDS=0
timeout 900 python -m pytest tests/test_layers_gpu.py tests/test_packnet_gpu.py tests/test_folded_gpu.py tests/test_recompose_gpu.py tests/test_step_gpu.py tests/test_graph_gpu.py -m gpu -q > ${O}_tests.log 2>&1; echo "gpu tests: rc $? in ${SECONDS}s"; grep -E "passed|failed|FAILED|Error" ${O}_tests.log | cut -c1-300 | head
for v in "" "PN_STENCIL_ASYNC=0"; do
  env $v timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-stock-torch > ${O}_bench.log 2> ${O}_bench.err
  echo "[$v] $(grep -o '"ms_per_step": [0-9.]*' ${O}_bench.log | head -1)"
done
timeout 300 python tools/step_profile.py --steps 2 --top 70 > ${O}_step_profile.log 2>&1; grep -E "device time|stencil|direct_copy|CUDAFunctor_add|frame" ${O}_step_profile.log | cut -c1-150
PN_STENCIL_ASYNC=0 timeout 300 python tools/step_profile.py --steps 2 --top 70 2>&1 | grep -E "device time|stencil" | cut -c1-150
