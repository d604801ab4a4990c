#!/bin/bash
# Round 2, final 1-GPU pass: whole suite, ncu capture of the tensor-core kernel, launch list, full bench line, reference arm.
set -u
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -q -m gpu --timeout 300 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
export_rep() { # name
    ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1.raw.csv 2>/dev/null
    ncu -i gpurun_out/$1.ncu-rep --page source --csv > gpurun_out/$1.source.csv 2>/dev/null
    rm -f gpurun_out/$1.ncu-rep
}
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:scan_tensor -s 1 -c 1 -f -o gpurun_out/r02_prof_tensor_100m python tools/prof_tensor.py 100000000 128 2 > gpurun_out/ncu_tensor.log 2>&1
echo "ncu tensor rc=$?"; tail -2 gpurun_out/ncu_tensor.log; export_rep r02_prof_tensor_100m
GSB_TC_DEBUG=1 timeout -s KILL 120 python tools/prof_tensor.py 100000000 128 2 > gpurun_out/r02_tensor_roles.log 2>&1; tail -8 gpurun_out/r02_tensor_roles.log
timeout -s KILL 300 python tools/tensor_try.py 100000000 > gpurun_out/r02_tensor_vs_sliced.log 2>&1; tail -13 gpurun_out/r02_tensor_vs_sliced.log
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_bench_1b.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-reference-cuda --no-full-verify --no-small-shard > gpurun_out/ncu_bench.log 2>&1
echo "ncu launches rc=$?"; grep -c "scan_topk\|sliced\|merge\|tensor" gpurun_out/r02_launches_bench_1b.csv
timeout -s KILL 1500 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
echo "bench rc=$?"; cat gpurun_out/r02_bench_n1.json; tail -5 gpurun_out/r02_bench_n1.err
timeout -s KILL 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err
echo "bench ref rc=$?"; cat gpurun_out/r02_bench_ref.json
du -sh gpurun_out
