#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
O=gpurun_out/r03d
timeout 300 python -m pytest tests/test_step_gpu.py tests/test_graph_gpu.py tests/test_optim_gpu.py -m gpu -q > ${O}_tests.log 2>&1; echo "tests rc $?"; tail -1 ${O}_tests.log | cut -c1-200
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-stock-torch > ${O}_bench.log 2> ${O}_bench.err; echo "$(grep -o '"ms_per_step": [0-9.]*' ${O}_bench.log | head -1) $(grep -c 'capture failed' ${O}_bench.err)"
