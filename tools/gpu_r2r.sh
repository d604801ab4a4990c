#!/bin/bash
# barrier-free tail: parity first (short timeouts), then the fixed-cost numbers with both forms
set -u
mkdir -p gpurun_out
timeout -s KILL 500 python -m pytest tests/test_gpu_parity.py -x -q -m gpu --timeout 200 -k "not full_size and not partition_invariance" > gpurun_out/pytest_tail.log 2>&1
echo "pytest rc=$?"; tail -6 gpurun_out/pytest_tail.log
for t in 1 0; do
echo "== GSB_TAIL=$t"; GSB_TAIL=$t timeout -s KILL 200 python tools/lat10m.py 2>&1 | grep "warps=16"
done
