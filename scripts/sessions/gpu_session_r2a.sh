#!/bin/bash
# r02a (1 GPU): first run of the resident latency kernel: functional check vs the oracle, ping floor, p50s,
# phase trace, then the GPU test-suite
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 300 python scripts/resident_check.py > gpurun_out/r02a_resident_check.txt 2>&1; echo "resident_check rc=$?"
tail -25 gpurun_out/r02a_resident_check.txt
YSM_TRACE=1 timeout 120 python scripts/latency_probe.py 360 1 > gpurun_out/r02a_trace_cfg1.txt 2>&1; echo "trace rc=$?"
tail -40 gpurun_out/r02a_trace_cfg1.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02a_pytest.txt 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r02a_pytest.txt
