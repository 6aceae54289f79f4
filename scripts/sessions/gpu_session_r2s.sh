#!/bin/bash
# r02s (1 GPU): in-stream per-kernel times, k_tile_stamp_lists vs k_tile_stamp
mkdir -p gpurun_out
timeout 300 python scripts/kernel_trace.py > gpurun_out/r02s_trace_new.txt 2>&1
YSM_TRACE_DEBUG=64 timeout 300 python scripts/kernel_trace.py > gpurun_out/r02s_trace_old.txt 2>&1
for f in new old; do echo "--- $f"; sed -n '/==== last call/,$p' gpurun_out/r02s_trace_$f.txt | head -40; done
