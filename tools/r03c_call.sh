#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
O=gpurun_out/r03c
timeout 600 python -m pytest tests/test_loss_gpu.py -m gpu -q > ${O}_tests.log 2>&1; echo "loss tests rc $?"; tail -2 ${O}_tests.log | cut -c1-200
for v in "PN_LOSS_MINB=2" "PN_LOSS_MINB=3"; do
  echo "[$v] $(env $v timeout 200 python tools/loss_only.py 2>&1 | grep 'graph replay')"
done
PN_LOSS_MINB=3 timeout 600 python -m pytest tests/test_loss_gpu.py -m gpu -q -k "grouped" > ${O}_tests3.log 2>&1; echo "loss tests (minb 3) rc $?"; tail -2 ${O}_tests3.log | cut -c1-200
for v in "PN_LOSS_MINB=2" "PN_LOSS_MINB=3"; do
  env $v timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-stock-torch > ${O}_bench.log 2> ${O}_bench.err
  echo "[$v] $(grep -o '"ms_per_step": [0-9.]*' ${O}_bench.log | head -1)"
done
