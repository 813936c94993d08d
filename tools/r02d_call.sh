#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
O=gpurun_out/r02d
timeout 900 python -m pytest tests/test_loss_gpu.py tests/test_step_gpu.py -m gpu -x -q > ${O}_tests_new.log 2>&1; echo "new tests: $?"; grep -v Warning ${O}_tests_new.log | tail -12
timeout 1500 python -m pytest tests -m gpu -q > ${O}_tests.log 2>&1; echo "gpu tier: $?"; tail -6 ${O}_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > ${O}_smoke.log 2>&1; echo "smoke: $?"; tail -2 ${O}_smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-stock-torch > ${O}_bench.log 2> ${O}_bench.err; echo "bench: $?"; cut -c1-600 ${O}_bench.log; tail -6 ${O}_bench.err
timeout 300 python tools/loss_only.py > ${O}_loss_only.txt 2>&1; tail -1 ${O}_loss_only.txt
PN_LOSS_FULLRES=1 timeout 300 python tools/loss_only.py > ${O}_loss_only_fullres.txt 2>&1; tail -1 ${O}_loss_only_fullres.txt
timeout 1500 bash tools/profile_r02.sh r02d > ${O}_profile.log 2>&1; tail -15 ${O}_profile.log
