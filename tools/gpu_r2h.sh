#!/bin/bash
# Round 2 re-entry: whole GPU suite on the restored tree, then tensor-core vs bit-sliced timings.
set -u
mkdir -p gpurun_out
timeout -s KILL 1800 python -m pytest tests -q -m gpu --timeout 900 --durations=15 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -25 gpurun_out/pytest_gpu.log
timeout -s KILL 900 python tools/tensor_try.py 100000000 > gpurun_out/tensor_try.log 2>&1
echo "tensor_try rc=$?"; tail -20 gpurun_out/tensor_try.log
