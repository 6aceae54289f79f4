#!/bin/bash
mkdir -p gpurun_out
YSM_TRACE=1 timeout 120 python scripts/latency_probe.py 360 1 > gpurun_out/r02m_trace_cfg1.txt 2>&1; echo "trace rc=$?"
tail -12 gpurun_out/r02m_trace_cfg1.txt | head -8
