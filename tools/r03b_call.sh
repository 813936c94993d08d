#!/bin/bash
# validation as the driver runs it + profiles of the final state
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
O=gpurun_out/r03b
SECONDS=0
timeout 1200 python -m pytest tests -m gpu -q > ${O}_tests.log 2>&1; echo "gpu tests: rc $? in ${SECONDS}s"; grep -E "passed|failed|FAILED|Error" ${O}_tests.log | cut -c1-300 | head
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
SECONDS=0
timeout 600 python bench.py --steps 20 --warmup 5 > ${O}_bench.log 2> ${O}_bench.err; echo "default bench rc $? in ${SECONDS}s"; cut -c1-300 ${O}_bench.log
env PN_WGRAD_STREAM=0 PN_PREFOLD_STREAM=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-stock-torch > ${O}_bench_nostreams.log 2> ${O}_bench_nostreams.err; echo "[no side streams] $(grep -o '"ms_per_step": [0-9.]*' ${O}_bench_nostreams.log | head -1)"
timeout 300 python tools/step_profile.py --steps 2 --top 70 > ${O}_step_profile.log 2>&1; head -3 ${O}_step_profile.log | tail -1
timeout 300 python bench.py --steps 10 --warmup 3 --batch 2 --height 384 --width 1280 --no-cpu-baseline --no-stock-torch > ${O}_bench_cfg3.log 2> ${O}_bench_cfg3.err; echo "cfg3: $(grep -o '"value": [0-9.]*' ${O}_bench_cfg3.log | head -1) $(grep -o '"ms_per_step": [0-9.]*' ${O}_bench_cfg3.log | head -1)"
NCU="ncu --clock-control none"
PN_CUDA_PROFILER=1 $NCU --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file ${O}_launches.csv python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-stock-torch > ${O}_launches_bench.log 2>&1; echo "launch list rc $?"
