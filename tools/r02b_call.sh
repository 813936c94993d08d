#!/bin/bash
# second GPU call of round 2: new defaults (fold, grouped loss, flat staging, GN tree, im2col, tiled unpack) + whole-step graph
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
O=gpurun_out/r02b
timeout 1500 python -m pytest tests -m gpu -x -q > ${O}_tests.log 2>&1; echo "gpu tier: $?"; tail -15 ${O}_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 > ${O}_bench.log 2> ${O}_bench.err; echo "bench: $?"; cut -c1-1500 ${O}_bench.log; tail -25 ${O}_bench.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-graph --no-cpu-baseline --no-stock-torch > ${O}_bench_nograph.log 2> ${O}_bench_nograph.err; echo "bench no graph: $?"; cut -c1-400 ${O}_bench_nograph.log
timeout 300 python tools/layer_table.py > ${O}_layer_table.txt 2>&1; tail -70 ${O}_layer_table.txt
