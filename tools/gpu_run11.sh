#!/bin/bash
set -u
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 300 -x -k "multi_query or repeat" > gpurun_out/pytest_batch.log 2>&1
echo "pytest batch rc=$?"; tail -15 gpurun_out/pytest_batch.log
timeout -s KILL 300 python tools/batch_bench.py 20000000 256 100 2>&1 | tail -4
timeout -s KILL 300 python tools/batch_bench.py 100000000 64 100 2>&1 | tail -4
