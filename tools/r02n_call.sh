#!/bin/bash
# ncu: weight-gradient kernel on the two worst big-map shapes (--set full), GroupNorm launches of one step (time + DRAM bytes)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
O=gpurun_out/r02n
NCU="ncu --clock-control none"
$NCU --set full --import-source on -k regex:"conv_wgrad_kernel" -s 2 -c 1 -f -o ${O}_wgrad_64_k3 python tools/conv_only.py 4 96 320 64 64 3 wgrad > ${O}_ncu.log 2>&1
$NCU --set full --import-source on -k regex:"conv_wgrad_kernel" -s 2 -c 1 -f -o ${O}_wgrad_136_k3 python tools/conv_only.py 4 192 640 136 64 3 wgrad >> ${O}_ncu.log 2>&1
PN_CUDA_PROFILER=1 $NCU --profile-from-start off -k regex:"gn_|split_bf16|stencil|head_|frame_|adam" --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv --log-file ${O}_gn_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-stock-torch > ${O}_gn_bench.log 2>&1
for f in ${O}_wgrad_64_k3 ${O}_wgrad_136_k3; do ncu -i $f.ncu-rep --page raw --csv > $f.csv 2>/dev/null; done
ls -la gpurun_out | tail -8
