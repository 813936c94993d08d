#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
O=gpurun_out/r02q
for mb in 0.001 4.5 8.5 16; do
  PN_GN_CLUSTER_MAX_MB=$mb timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-stock-torch > ${O}_bench_$mb.log 2> ${O}_bench_$mb.err
  echo "max_mb=$mb: $(grep -o '"ms_per_step": [0-9.]*' ${O}_bench_$mb.log | head -1)"
done
