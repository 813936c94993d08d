#!/bin/bash
# cluster GroupNorm kernels + float4 apply kernel + balanced wgrad tap groups: parity tests, CUPTI breakdown, bench
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
O=gpurun_out/r02o
SECONDS=0
timeout 900 python -m pytest tests/test_recompose_gpu.py tests/test_layers_gpu.py tests/test_packnet_gpu.py tests/test_folded_gpu.py tests/test_step_gpu.py tests/test_graph_gpu.py -m gpu -q > ${O}_tests.log 2>&1; echo "gpu tests: rc $? in ${SECONDS}s"; grep -E "passed|failed|FAILED|Error" ${O}_tests.log | cut -c1-300 | head -20
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-stock-torch > ${O}_bench.log 2> ${O}_bench.err; echo "bench rc $?"; cut -c1-400 ${O}_bench.log
grep -E "timed region|e2e region|fail" ${O}_bench.err | cut -c1-300
timeout 300 python tools/step_profile.py --steps 2 --top 40 > ${O}_step_profile.log 2>&1; echo "step_profile rc $?"; head -45 ${O}_step_profile.log | cut -c1-170
