#!/bin/bash
# Multi-query kernels on every row width: parity tests, then bit-sliced vs POPC timings.
set -u
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -q -m gpu -k "sliced or multi_query or batch_kernel_choice" --timeout 600 2>&1 | tail -3
timeout -s KILL 400 python tools/batch_width_sweep.py > gpurun_out/batch_width_sweep.log 2>&1
echo "sweep rc=$?"; tail -8 gpurun_out/batch_width_sweep.log
