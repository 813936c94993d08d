#!/bin/bash
# per-layer table of the current default path, rest of the GPU tier after the step test, BASELINE configs[2] shape (B=2, 384x1280)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
O=gpurun_out/r02l
SECONDS=0
timeout 600 python -m pytest tests/test_step_gpu.py tests/test_ddp_gpu.py tests/test_recompose_gpu.py tests/test_packnet_gpu.py -m gpu -q -s > ${O}_tests.log 2>&1; echo "gpu tests (step..): rc $? in ${SECONDS}s"; grep -E "passed|failed|FAILED|smallest fractions" ${O}_tests.log | cut -c1-400
SECONDS=0
timeout 300 python tools/layer_table.py > ${O}_layer_table.log 2>&1; echo "layer_table rc $? in ${SECONDS}s"; head -90 ${O}_layer_table.log | cut -c1-200
SECONDS=0
timeout 600 python bench.py --steps 10 --warmup 3 --batch 2 --height 384 --width 1280 --no-cpu-baseline > ${O}_bench_cfg3.log 2> ${O}_bench_cfg3.err; echo "bench cfg3 rc $? in ${SECONDS}s"; cut -c1-700 ${O}_bench_cfg3.log
grep -E "timed region|e2e region|enqueue|captured|stock|fail" ${O}_bench_cfg3.err | cut -c1-400
