#!/bin/bash
# r02k (1 GPU): phase trace of the resident kernel with timestamps inside the sweep and the stamping
mkdir -p gpurun_out
YSM_TRACE=1 timeout 120 python scripts/latency_probe.py 360 1 > gpurun_out/r02k_trace_cfg1.txt 2>&1; echo "trace rc=$?"
tail -34 gpurun_out/r02k_trace_cfg1.txt
