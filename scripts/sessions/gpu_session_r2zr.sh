#!/bin/bash
# r02zr (1 GPU): resident collect skips warps past the cell list: resident tests, latency probes x3 per shape
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_resident.py tests/test_gpu_reference_slam.py -m gpu -x -q 2>&1 | tail -2
for S in "360 1" "720 10" "360 1" "720 10" "360 1" "720 10"; do timeout 120 python scripts/latency_probe.py $S 2>&1 | grep "Wrapper.match_scan"; done
YSM_TRACE=1 timeout 120 python scripts/latency_probe.py 360 1 2>&1 | grep "workers collect\|workers stamp\|resident: results" | tail -3
