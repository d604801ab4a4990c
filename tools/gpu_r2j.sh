#!/bin/bash
# tensor-core kernel: timing experiments (fault 2: no expander stores, no candidates; 3: no candidates; 4: no stores)
set -u
mkdir -p gpurun_out
for f in 0 2 3 4; do
GSB_TC_FAULT=$f timeout -s KILL 600 python tools/prof_tensor.py 100000000 128 3 > gpurun_out/tensor_time_$f.log 2>&1
echo "fault $f rc=$?"; tail -3 gpurun_out/tensor_time_$f.log | head -2
GSB_TC_DEBUG=1 GSB_TC_FAULT=$f timeout -s KILL 600 python tools/prof_tensor.py 100000000 128 2 2>&1 | tail -8 | head -7
done
timeout -s KILL 600 python tools/lat10m.py > gpurun_out/lat10m.log 2>&1; grep "warps=16" gpurun_out/lat10m.log
