#!/bin/bash
# validation as the driver runs it (whole GPU tier, smoke, default bench) + frame K split A/B + ncu launch list / dominant conv capture + config 3
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
O=gpurun_out/r02x
SECONDS=0
timeout 1200 python -m pytest tests -m gpu -q > ${O}_tests.log 2>&1; echo "gpu tests: rc $? in ${SECONDS}s"; grep -E "passed|failed|FAILED|Error" ${O}_tests.log | cut -c1-300 | head
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for v in "" "PN_FRAME_KSPLIT=0"; do
  env $v timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-stock-torch > ${O}_bench_ab.log 2> ${O}_bench_ab.err
  echo "[$v] $(grep -o '"ms_per_step": [0-9.]*' ${O}_bench_ab.log | head -1)"
done
SECONDS=0
timeout 600 python bench.py --steps 20 --warmup 5 > ${O}_bench.log 2> ${O}_bench.err; echo "default bench rc $? in ${SECONDS}s"; cut -c1-300 ${O}_bench.log
timeout 300 python tools/step_profile.py --steps 2 --top 70 > ${O}_step_profile.log 2>&1; grep -E "device time|frame" ${O}_step_profile.log | cut -c1-150
NCU="ncu --clock-control none"
PN_CUDA_PROFILER=1 $NCU --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file ${O}_launches.csv python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-stock-torch > ${O}_launches_bench.log 2>&1
$NCU --set full --import-source on -k regex:"conv_igemm_kernel" -s 2 -c 1 -f -o ${O}_conv_pack1f python tools/conv_only.py 4 96 320 256 64 7 fwd > ${O}_conv_ncu.log 2>&1
ncu -i ${O}_conv_pack1f.ncu-rep --page raw --csv > ${O}_conv_pack1f.csv 2>/dev/null
timeout 300 python bench.py --steps 10 --warmup 3 --batch 2 --height 384 --width 1280 --no-cpu-baseline --no-stock-torch > ${O}_bench_cfg3.log 2> ${O}_bench_cfg3.err; echo "cfg3: $(grep -o '"value": [0-9.]*' ${O}_bench_cfg3.log | head -1) $(grep -o '"ms_per_step": [0-9.]*' ${O}_bench_cfg3.log | head -1)"
rm -f ${O}_conv_pack1f.ncu-rep
