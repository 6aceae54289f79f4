#!/bin/bash
# r02h (1 GPU): resident kernel + device-resident scan store (content tags), tagged doorbell lines
mkdir -p gpurun_out
timeout 300 python scripts/resident_check.py > gpurun_out/r02h_resident_check.txt 2>&1; echo "resident_check rc=$?"
tail -22 gpurun_out/r02h_resident_check.txt
YSM_TRACE=1 timeout 120 python scripts/latency_probe.py 360 1 > gpurun_out/r02h_trace_cfg1.txt 2>&1; echo "trace rc=$?"
tail -24 gpurun_out/r02h_trace_cfg1.txt
YSM_TRACE=1 timeout 120 python scripts/latency_probe.py 720 10 > gpurun_out/r02h_trace_cfg2.txt 2>&1; echo "trace rc=$?"
tail -24 gpurun_out/r02h_trace_cfg2.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02h_pytest.txt 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r02h_pytest.txt
