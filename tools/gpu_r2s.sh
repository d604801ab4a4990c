#!/bin/bash
set -u
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu --timeout 200 -k "select_with_and_without or tuning_knobs or adversarial or large_k or ties" > gpurun_out/pytest_tail.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/pytest_tail.log
for t in 1 0; do
echo "== GSB_TAIL=$t"; GSB_TAIL=$t timeout -s KILL 200 python tools/lat10m.py 2>&1 | grep "warps=16"
done
echo "== GSB_TAIL=1 GSB_SHARE_HIST=1"; GSB_TAIL=1 GSB_SHARE_HIST=1 timeout -s KILL 200 python tools/lat10m.py 2>&1 | grep "warps=16"
