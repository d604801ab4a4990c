#!/bin/bash
# BASELINE config 5 (1 B rows sharded, 1024 queries, top-100) at N GPUs: bit-sliced kernel, POPC kernel for comparison.
set -u
N=${1:-8}
mkdir -p gpurun_out
for mode in 1 2; do
GSB_BATCH_KERNEL=$mode timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tools/bench_batch_dist.py --rows 1000000000 --queries 1024 --k 100 > gpurun_out/batch_dist_n${N}_mode$mode.json 2> gpurun_out/batch_dist_n${N}_mode$mode.err
echo "batch dist n=$N mode=$mode rc=$?"; tail -1 gpurun_out/batch_dist_n${N}_mode$mode.json; grep -v "OMP_NUM\|\*\*\*\*\|NCCL version" gpurun_out/batch_dist_n${N}_mode$mode.err | tail -3
done
