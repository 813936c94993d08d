#!/bin/bash
# third GPU call of round 2: data gradient from the forward tiles (BT), FlatAdam with stored weights
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
O=gpurun_out/r02c
timeout 600 python -m pytest tests/test_layers_gpu.py tests/test_optim_gpu.py tests/test_step_gpu.py -m gpu -x -q -s > ${O}_tests_new.log 2>&1; echo "new tests: $?"; grep -v Warning ${O}_tests_new.log | tail -30
timeout 1500 python -m pytest tests -m gpu -q > ${O}_tests.log 2>&1; echo "gpu tier: $?"; tail -8 ${O}_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-stock-torch > ${O}_bench.log 2> ${O}_bench.err; echo "bench: $?"; cut -c1-700 ${O}_bench.log; tail -12 ${O}_bench.err
timeout 300 python bench.py --steps 20 --warmup 5 --torch-adam --no-cpu-baseline --no-stock-torch > ${O}_bench_torchadam.log 2> ${O}_bench_torchadam.err; echo "bench torch adam: $?"; cut -c1-400 ${O}_bench_torchadam.log
timeout 300 python tools/step_profile.py > ${O}_step_breakdown.txt 2>&1; head -45 ${O}_step_breakdown.txt
