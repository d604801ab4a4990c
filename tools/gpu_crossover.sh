#!/bin/bash
# Bit-sliced kernel: parity tests, small-batch cross-over against the POPC kernel, the 1024-query bench.
set -u
timeout -s KILL 900 python -m pytest tests -q -m gpu -k "sliced or multi_query or batch_kernel_choice or full_size" --timeout 600 2>&1 | tail -3
timeout -s KILL 300 python tools/crossover.py 100000000 2>&1 | grep rows=
timeout -s KILL 300 python tools/batch_bench.py 100000000 1024 100 fast 2>&1 | grep "bit-sliced"
