#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
O=gpurun_out/r02e
timeout 900 python -m pytest tests/test_layers_gpu.py tests/test_step_gpu.py tests/test_packnet_gpu.py -m gpu -x -q > ${O}_tests_new.log 2>&1; echo "new tests: $?"; grep -v Warning ${O}_tests_new.log | tail -8
timeout 1500 python -m pytest tests -m gpu -q > ${O}_tests.log 2>&1; echo "gpu tier: $?"; tail -6 ${O}_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-stock-torch > ${O}_bench.log 2> ${O}_bench.err; echo "bench: $?"; cut -c1-500 ${O}_bench.log; tail -4 ${O}_bench.err
PN_GN_EMIT_SPLIT=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-stock-torch > ${O}_bench_noemit.log 2> ${O}_bench_noemit.err; echo "bench (no emit): $?"; cut -c1-300 ${O}_bench_noemit.log
timeout 300 python tools/posenet_bench.py > ${O}_posenet.txt 2>&1; cat ${O}_posenet.txt | grep -v Warn
timeout 300 python tools/step_profile.py > ${O}_step_breakdown.txt 2>&1; head -30 ${O}_step_breakdown.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-stock-torch --height 384 --width 1280 --batch 2 > ${O}_bench_384.log 2> ${O}_bench_384.err; echo "bench 384x1280: $?"; cut -c1-400 ${O}_bench_384.log; tail -3 ${O}_bench_384.err
