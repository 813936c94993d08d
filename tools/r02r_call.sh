#!/bin/bash
# fused loss launch: tests on the GPU, event timing, ncu --set full with source counters
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
O=gpurun_out/r02r
timeout 600 python -m pytest tests/test_loss_gpu.py -m gpu -q > ${O}_tests.log 2>&1; echo "loss tests rc $?"; tail -3 ${O}_tests.log | cut -c1-200
timeout 200 python tools/loss_only.py > ${O}_loss_only.log 2>&1; tail -6 ${O}_loss_only.log
ncu --clock-control none --set full --import-source on -k regex:"loss_group_kernel" -s 2 -c 1 -f -o ${O}_loss python tools/loss_only.py > ${O}_ncu.log 2>&1
ncu -i ${O}_loss.ncu-rep --page raw --csv > ${O}_loss_raw.csv 2>/dev/null
ncu -i ${O}_loss.ncu-rep --page source --csv > ${O}_loss_source.csv 2>/dev/null
ls -la gpurun_out | tail -5
