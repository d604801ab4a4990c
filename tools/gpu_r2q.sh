#!/bin/bash
set -u
for f in 8 5 0; do
echo "fault $f"; GSB_TC_FAULT=$f timeout -s KILL 120 python tools/prof_tensor.py 100000000 128 3 2>&1 | tail -3 | head -2
done
