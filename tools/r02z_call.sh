#!/bin/bash
# side-stream weight gradients: parity / graph / step / optimizer tests, A/B, breakdown
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
O=gpurun_out/r02z
SECONDS=0
timeout 900 python -m pytest tests/test_layers_gpu.py tests/test_packnet_gpu.py tests/test_folded_gpu.py tests/test_step_gpu.py tests/test_graph_gpu.py tests/test_optim_gpu.py tests/test_recompose_gpu.py -m gpu -q > ${O}_tests.log 2>&1; echo "gpu tests: rc $? in ${SECONDS}s"; grep -E "passed|failed|FAILED|Error" ${O}_tests.log | cut -c1-300 | head
for v in "" "PN_WGRAD_STREAM=0" ""; do
  env $v timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-stock-torch > ${O}_bench_ab.log 2> ${O}_bench_ab.err
  echo "[$v] $(grep -o '"ms_per_step": [0-9.]*' ${O}_bench_ab.log | head -1) $(grep -c 'capture failed' ${O}_bench_ab.err) $(grep -o '"loss": [0-9.]*' ${O}_bench_ab.log | head -1)"
done
timeout 300 python bench.py --steps 10 --warmup 3 --no-graph --no-cpu-baseline --no-stock-torch > ${O}_bench_eager.log 2> ${O}_bench_eager.err; echo "eager: $(grep -o '"ms_per_step": [0-9.]*' ${O}_bench_eager.log | head -1)"
