#!/bin/bash
set -u
mkdir -p gpurun_out
for f in 2 3 4 0; do
GSB_TC_FAULT=$f timeout -s KILL 600 python tools/prof_tensor.py 100000000 128 3 > gpurun_out/tensor_time_$f.log 2>&1
echo "fault $f rc=$?"; tail -3 gpurun_out/tensor_time_$f.log | head -2
done
GSB_TC_VARIANT=1 timeout -s KILL 600 python tools/prof_tensor.py 100000000 128 3 > gpurun_out/tensor_time_v1.log 2>&1
echo "variant 1 rc=$?"; tail -3 gpurun_out/tensor_time_v1.log | head -2
timeout -s KILL 1500 python -m pytest tests/test_gpu_tensor.py -x -q -m gpu --timeout 600 > gpurun_out/pytest_tensor.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/pytest_tensor.log
