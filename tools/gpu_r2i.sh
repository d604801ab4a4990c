#!/bin/bash
# tensor-core kernel: role timers and one full ncu capture (exported to CSV on the box)
set -u
mkdir -p gpurun_out
GSB_TC_DEBUG=1 timeout -s KILL 600 python tools/prof_tensor.py 100000000 128 3 > gpurun_out/tensor_dbg.log 2>&1
echo "dbg rc=$?"; tail -12 gpurun_out/tensor_dbg.log
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:scan_tensor -s 1 -c 1 -f -o gpurun_out/r02_prof_tensor_100m python tools/prof_tensor.py 100000000 128 2 > gpurun_out/ncu_tensor.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_tensor.log
for p in raw details source; do ncu -i gpurun_out/r02_prof_tensor_100m.ncu-rep --page $p --csv > gpurun_out/r02_prof_tensor_100m.$p.csv 2>/dev/null; done
rm -f gpurun_out/r02_prof_tensor_100m.ncu-rep
du -sh gpurun_out
