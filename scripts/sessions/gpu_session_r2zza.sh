#!/bin/bash
# r02zza (1 GPU): k_find_valid with the block-wise trigger chain: parity suite, in-stream kernel trace
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 200 python scripts/kernel_trace.py > gpurun_out/r02zza_trace.txt 2>&1
sed -n '/==== last call/,$p' gpurun_out/r02zza_trace.txt | grep "k_find_valid\|k_tile_stamp"
