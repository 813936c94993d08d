#!/bin/bash
# baseline call of the re-entered session: whole GPU tier, default bench (N=1), CUPTI kernel breakdown of a step
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
O=gpurun_out/r02k
SECONDS=0
timeout 900 python -m pytest tests -m gpu -q -x > ${O}_tests.log 2>&1; echo "gpu tests: rc $? in ${SECONDS}s"; tail -4 ${O}_tests.log | cut -c1-300
SECONDS=0
timeout 600 python bench.py --steps 20 --warmup 5 > ${O}_bench.log 2> ${O}_bench.err; echo "bench rc $? in ${SECONDS}s"; cut -c1-1500 ${O}_bench.log
grep -E "timed region|e2e region|enqueue|captured|stock" ${O}_bench.err | cut -c1-400
SECONDS=0
timeout 300 python tools/step_profile.py --steps 2 --top 70 > ${O}_step_profile.log 2>&1; echo "step_profile rc $? in ${SECONDS}s"; head -75 ${O}_step_profile.log | cut -c1-200
