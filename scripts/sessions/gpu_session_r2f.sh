#!/bin/bash
# r02f (1 GPU): SM clock while the resident kernel serves requests
mkdir -p gpurun_out
YSM_TRACE=1 timeout 120 python scripts/latency_probe.py 360 1 > gpurun_out/r02f_trace_cfg1.txt 2>&1; echo "trace rc=$?"
grep "SM clock" gpurun_out/r02f_trace_cfg1.txt | tail -4
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
