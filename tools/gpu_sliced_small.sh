#!/bin/bash
# where does the fixed cost of a bit-sliced pass go?  ncu source view of a 150 k-row, 1024-query launch
set -u
mkdir -p gpurun_out
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:scan_sliced -s 1 -c 1 -f -o gpurun_out/sliced_small python tools/prof_sliced.py 150000 1024 2 > gpurun_out/ncu_sliced_small.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_sliced_small.log
ncu -i gpurun_out/sliced_small.ncu-rep --page raw --csv > gpurun_out/sliced_small.raw.csv 2>/dev/null
ncu -i gpurun_out/sliced_small.ncu-rep --page source --csv > gpurun_out/sliced_small.source.csv 2>/dev/null
rm -f gpurun_out/sliced_small.ncu-rep
