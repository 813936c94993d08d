#!/bin/bash
# loss gradient program after the restructuring: tests, timing (fused / two launches / tile), ncu totals
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
O=gpurun_out/r02s
timeout 600 python -m pytest tests/test_loss_gpu.py -m gpu -q > ${O}_tests.log 2>&1; echo "loss tests rc $?"; tail -3 ${O}_tests.log | cut -c1-200
timeout 200 python tools/loss_only.py 2>&1 | grep program
PN_LOSS_FUSED=0 timeout 200 python tools/loss_only.py 2>&1 | grep program
ncu --clock-control none --set full --import-source on -k regex:"loss_group_kernel" -s 2 -c 1 -f -o ${O}_loss python tools/loss_only.py > ${O}_ncu.log 2>&1
ncu -i ${O}_loss.ncu-rep --page raw --csv > ${O}_loss_raw.csv 2>/dev/null
ncu -i ${O}_loss.ncu-rep --page source --csv > ${O}_loss_source.csv 2>/dev/null
python - <<'PY'
import csv
r=list(csv.reader([l for l in open('gpurun_out/r02s_loss_raw.csv',newline='') if not l.startswith('==')]))
d=dict(zip(r[0],r[2]))
for k in ('gpu__time_duration.sum','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','dram__bytes_read.sum','launch__grid_size','sm__warps_active.avg.pct_of_peak_sustained_active'): print(k,d.get(k))
PY
