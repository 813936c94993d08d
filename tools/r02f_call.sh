#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
O=gpurun_out/r02f
timeout 900 python -m pytest tests/test_step_gpu.py tests/test_graph_gpu.py -m gpu -q -s > ${O}_tests_new.log 2>&1; echo "new tests: $?"; grep -v Warning ${O}_tests_new.log | grep -E "largest|   [0-9]\.|passed|failed|FAILED|Error|graph \[" | head -40
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-stock-torch > ${O}_bench.log 2> ${O}_bench.err; echo "bench: $?"; cut -c1-300 ${O}_bench.log; tail -4 ${O}_bench.err
PN_POSE_STREAM=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-stock-torch > ${O}_bench_nostream.log 2> ${O}_bench_nostream.err; echo "bench (PoseNet on the main stream): $?"; cut -c1-300 ${O}_bench_nostream.log
