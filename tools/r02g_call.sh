#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
O=gpurun_out/r02g
for ps in 0 1; do
  PN_POSE_STREAM=$ps timeout 900 python -m pytest tests/test_step_gpu.py -m gpu -q -s > ${O}_tests_ps$ps.log 2>&1; echo "PN_POSE_STREAM=$ps: $?"
  grep -v Warning ${O}_tests_ps$ps.log | grep -E "largest|   [0-9]\.[0-9]+ |passed|failed|FAILED|smoothness_loss'|step [01] " | head -30
done
