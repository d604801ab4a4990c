#!/bin/bash
set -u
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -q -m gpu --timeout 300 -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log
timeout -s KILL 600 python tools/bench_batch_dist.py --rows 125000000 --queries 512 --k 100 2>&1 | tail -3
