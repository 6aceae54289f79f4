#!/bin/bash
mkdir -p gpurun_out
BENCH="python bench.py --steps 1 --warmup 1 --matches 20000 --no-latency --no-extras --no-cpu"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_find_valid -s 30 -c 1 -f -o gpurun_out/r02zj_k_find_valid $BENCH > gpurun_out/r02zj_k_find_valid.log 2>&1; echo "rc=$?"
