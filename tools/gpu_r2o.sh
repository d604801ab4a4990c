#!/bin/bash
# tensor kernel with / without setmaxnreg: short timeouts (a hang must not eat the budget)
set -u
mkdir -p gpurun_out
echo "== with setmaxnreg"
timeout -s KILL 60 python tools/tensor_try.py 2>&1 | tail -4
cp gpusimilarity_b200/libgpusim_b200.so /tmp/lib_regs.so
cp build/libgpusim_b200_noregs.so gpusimilarity_b200/libgpusim_b200.so
echo "== without setmaxnreg"
timeout -s KILL 60 python tools/tensor_try.py 2>&1 | tail -4
timeout -s KILL 120 python tools/prof_tensor.py 100000000 128 3 2>&1 | tail -3 | head -2
GSB_TC_DEBUG=1 timeout -s KILL 120 python tools/prof_tensor.py 100000000 128 2 2>&1 | tail -8 | head -7
