#!/bin/bash
# r02zp (1 GPU): resident-kernel phase trace at HEAD, cfg-1 shape and cfg-2 shape
mkdir -p gpurun_out
YSM_TRACE=1 timeout 120 python scripts/latency_probe.py 360 1 > gpurun_out/r02zp_trace_cfg1.txt 2>&1; echo "rc=$?"
YSM_TRACE=1 timeout 120 python scripts/latency_probe.py 720 10 > gpurun_out/r02zp_trace_cfg2.txt 2>&1; echo "rc=$?"
grep "p50" gpurun_out/r02zp_trace_cfg1.txt gpurun_out/r02zp_trace_cfg2.txt
tail -34 gpurun_out/r02zp_trace_cfg2.txt | head -32
