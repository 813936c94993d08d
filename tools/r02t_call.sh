#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
O=gpurun_out/r02t
timeout 200 python tools/loss_only.py > ${O}_loss_only.log 2>&1; grep -v Warn ${O}_loss_only.log | tail -12
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-stock-torch > ${O}_bench.log 2> ${O}_bench.err; echo "bench rc $?"; cut -c1-330 ${O}_bench.log
grep -E "timed region|e2e region|fail" ${O}_bench.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02t_bench.log').read().strip().splitlines()[-1])
print(json.dumps(d.get('roofline_loss'))[:900])
PY
