#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
O=gpurun_out/r03a
SECONDS=0
timeout 900 python -m pytest tests/test_packnet_gpu.py tests/test_folded_gpu.py tests/test_step_gpu.py tests/test_graph_gpu.py tests/test_optim_gpu.py -m gpu -q > ${O}_tests.log 2>&1; echo "gpu tests: rc $? in ${SECONDS}s"; grep -E "passed|failed|FAILED|Error" ${O}_tests.log | cut -c1-300 | head
for v in "" "" ; do
  env $v timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-stock-torch > ${O}_bench_ab.log 2> ${O}_bench_ab.err
  echo "[$v] $(grep -o '"ms_per_step": [0-9.]*' ${O}_bench_ab.log | head -1) $(grep -c 'capture failed' ${O}_bench_ab.err)"
done
timeout 300 python tools/step_profile.py --steps 2 --top 30 > ${O}_step_profile.log 2>&1; head -34 ${O}_step_profile.log | cut -c1-150
