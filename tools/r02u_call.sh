#!/bin/bash
# two GPUs: NCCL gradient-equality test, bench N=2 with the overlapped buckets and with one trailing all-reduce, reference arm under torchrun
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
O=gpurun_out/r02u
timeout 300 python -m pytest tests/test_ddp_gpu.py -m gpu -q -s > ${O}_ddp_test.log 2>&1; echo "ddp test: $?"; grep -E "DDP |passed|failed|Error" ${O}_ddp_test.log | head -5
for flags in "" "--no-overlap"; do
  tag=$(echo "n2$flags" | tr -d ' -')
  SECONDS=0
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 $flags > ${O}_bench_${tag}.log 2> ${O}_bench_${tag}.err
  echo "bench N=2 [$flags]: rc $? in ${SECONDS}s $(cut -c1-330 ${O}_bench_${tag}.log)"; grep -E "capture|graph|timed region" ${O}_bench_${tag}.err | tail -3
done
