#!/bin/bash
# wave-aware tile / split plan of the convolution kernels: conv + network parity tests, per-layer table, bench
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
O=gpurun_out/r02m
SECONDS=0
timeout 900 python -m pytest tests/test_layers_gpu.py tests/test_packnet_gpu.py tests/test_folded_gpu.py tests/test_step_gpu.py -m gpu -q -x > ${O}_tests.log 2>&1; echo "gpu tests: rc $? in ${SECONDS}s"; grep -E "passed|failed|FAILED" ${O}_tests.log | cut -c1-400
timeout 300 python tools/layer_table.py > ${O}_layer_table.log 2>&1; echo "layer_table rc $?"; head -60 ${O}_layer_table.log | cut -c1-160
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-stock-torch > ${O}_bench.log 2> ${O}_bench.err; echo "bench rc $?"; cut -c1-400 ${O}_bench.log
grep -E "timed region|e2e region|fail" ${O}_bench.err | cut -c1-300
