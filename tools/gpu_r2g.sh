#!/bin/bash
# Round 2: ncu --set full of the tensor-core multi-query kernel (100 M rows, 128 queries).
set -u
mkdir -p gpurun_out
export_rep() { # name
    ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1.raw.csv 2>/dev/null
    ncu -i gpurun_out/$1.ncu-rep --page details --csv > gpurun_out/$1.details.csv 2>/dev/null
    ncu -i gpurun_out/$1.ncu-rep --page source --csv > gpurun_out/$1.source.csv 2>/dev/null
    rm -f gpurun_out/$1.ncu-rep
}
timeout -s KILL 600 python tools/prof_tensor.py 100000000 128 3 > gpurun_out/tensor_time.log 2>&1
echo "time rc=$?"; cat gpurun_out/tensor_time.log
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:scan_tensor -s 1 -c 1 -f -o gpurun_out/r02_prof_tensor_100m python tools/prof_tensor.py 100000000 128 2 > gpurun_out/ncu_tensor.log 2>&1
echo "ncu tensor rc=$?"; tail -3 gpurun_out/ncu_tensor.log; export_rep r02_prof_tensor_100m
