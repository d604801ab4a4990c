#!/bin/bash
set -u
mkdir -p gpurun_out
timeout -s KILL 600 python tools/daemon_bench.py 2000000 16 64 20 > gpurun_out/r02_daemon.log 2>&1; echo rc=$?; cat gpurun_out/r02_daemon.log
