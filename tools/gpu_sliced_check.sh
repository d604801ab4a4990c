#!/bin/bash
# bit-sliced + tensor-core kernels: parity tests, fixed cost of a pass, 100 M-row pass
set -u
mkdir -p gpurun_out
timeout -s KILL 500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tensor.py -x -q -m gpu --timeout 200 -k "sliced or tensor or multi_query or batch" > gpurun_out/pytest_sliced.log 2>&1
echo "pytest rc=$?"; tail -2 gpurun_out/pytest_sliced.log
timeout -s KILL 300 python tools/batch_fixed_cost.py 2>&1 | tail -6
timeout -s KILL 300 python tools/prof_sliced.py 100000000 1024 1 > /dev/null 2>&1
timeout -s KILL 300 python tools/tensor_try.py 100000000 2>&1 | grep "nq=1024\|PARITY"
