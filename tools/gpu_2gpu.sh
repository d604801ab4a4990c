#!/bin/bash
# 2-GPU pass: GPU tests (incl. the single-GPU two-rank fused test), bench N=2 (fused exchange; the line carries
# the multi_query pass), racecheck of the bit-sliced kernel.
set -u
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -q -m gpu --timeout 300 -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
echo "bench n2 rc=$?"; cat gpurun_out/bench_n2.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e'], d['gpu_launches'], d['config']['parallelism'], d['verified'], d.get('multi_query'))"; grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/bench_n2.err | tail -3
timeout -s KILL 600 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests -q -m gpu -k "sliced_kernel_sizes and 1025" --timeout 500 > gpurun_out/racecheck_sliced.log 2>&1
echo "racecheck rc=$?"; tail -3 gpurun_out/racecheck_sliced.log
