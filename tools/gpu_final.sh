#!/bin/bash
# Round-end style validation: smoke, the whole GPU suite, both bench arms at N=1, sanitizer passes on the
# bit-sliced kernel.
set -u
mkdir -p gpurun_out
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout -s KILL 1500 python -m pytest tests -q -m gpu --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu.log
timeout -s KILL 600 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
echo "bench ref rc=$?"; cut -c1-400 gpurun_out/bench_ref.json
timeout -s KILL 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "bench n1 rc=$?"; cat gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
timeout -s KILL 600 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests -q -m gpu -k "sliced_kernel_sizes and (1025 or 37889)" --timeout 500 > gpurun_out/memcheck_sliced.log 2>&1
echo "memcheck rc=$?"; tail -4 gpurun_out/memcheck_sliced.log
timeout -s KILL 600 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests -q -m gpu -k "sliced_kernel_sizes and 1025" --timeout 500 > gpurun_out/racecheck_sliced.log 2>&1
echo "racecheck rc=$?"; grep -c "Race reported\|hazard" gpurun_out/racecheck_sliced.log; tail -4 gpurun_out/racecheck_sliced.log
