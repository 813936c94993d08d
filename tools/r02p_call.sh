#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
O=gpurun_out/r02p
timeout 300 python tools/gn_bench.py 0 1 > ${O}_gn_bench.log 2>&1; cat ${O}_gn_bench.log
PN_GN_CLUSTER_CL=16 timeout 300 python tools/gn_bench.py 1 > ${O}_gn_bench16.log 2>&1; cat ${O}_gn_bench16.log
