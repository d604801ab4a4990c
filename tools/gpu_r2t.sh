#!/bin/bash
set -u
for t in 0 1; do echo "== GSB_TAIL=$t"; GSB_TAIL=$t timeout -s KILL 200 python tools/dbg_times.py 10000000 2>&1 | tail -12; done
