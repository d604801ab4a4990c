#!/bin/bash
# 8-GPU pass as the driver runs it: reference arm under torchrun (rank 0 works), then bench.py at N=8 (fused
# exchange; the line carries the multi_query pass = BASELINE configs[4]).
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 8 --steps 3 --warmup 3 > gpurun_out/bench_ref_n8.json 2> gpurun_out/bench_ref_n8.err
echo "reference arm n=8 rc=$?"; cut -c1-200 gpurun_out/bench_ref_n8.json
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 50 --warmup 5 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err
echo "bench n=8 rc=$?"; python -c "import sys,json; d=json.loads(open('gpurun_out/bench_n8.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['gpu_launches'], d['roofline']['kernel_ms'], d['config']['parallelism'], d['verified'], d['clocks'], d.get('multi_query'))"; grep -v "OMP_NUM\|\*\*\*\*\|NCCL version" gpurun_out/bench_n8.err | tail -5
