#!/bin/bash
set -u
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -q -m gpu --timeout 300 -x -k "multi_device or adapter or server" > gpurun_out/pytest_gpu2.log 2>&1
echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu2.log
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/bench_batch_dist.py --rows 250000000 --queries 512 --k 100 2>&1 | grep -v "OMP_NUM\|\*\*\*\*" | tail -3
