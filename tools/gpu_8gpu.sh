#!/bin/bash
# 8-GPU pass: bench at N=8 and N=4 (fused exchange), N=8 NCCL all-gather for comparison.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
for cfg in "8 1" "4 1" "8 0"; do
set -- $cfg
GSB_FUSED_EXCHANGE=$2 timeout -s KILL 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $1 --steps 50 --warmup 5 > gpurun_out/bench_n$1_fused$2.json 2> gpurun_out/bench_n$1_fused$2.err
echo "bench n=$1 fused=$2 rc=$?"; python -c "import sys,json; d=json.loads(open('gpurun_out/bench_n$1_fused$2.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['gpu_launches'], d['roofline']['kernel_ms'], d['config']['parallelism'], d['verified'], d['clocks'])"; grep -v "OMP_NUM\|\*\*\*\*\|NCCL version" gpurun_out/bench_n$1_fused$2.err | tail -5
done
