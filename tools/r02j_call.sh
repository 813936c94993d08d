#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
O=gpurun_out/r02j
timeout 300 python -m pytest tests/test_ddp_gpu.py -m gpu -q -s > ${O}_ddp_test.log 2>&1; echo "ddp test: $?"; grep -E "DDP |passed|failed|Error" ${O}_ddp_test.log | head -5
SECONDS=0
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > ${O}_bench_n2.log 2> ${O}_bench_n2.err
echo "bench N=2: rc $? in ${SECONDS}s $(cut -c1-330 ${O}_bench_n2.log)"; grep -E "capture|graph|timed region" ${O}_bench_n2.err | tail -3
