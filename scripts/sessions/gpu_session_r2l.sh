#!/bin/bash
# r02l (1 GPU): resident kernel with host-made penalty tables, winner record, covariance overlapped with the fine items
mkdir -p gpurun_out
timeout 300 python scripts/resident_check.py > gpurun_out/r02l_resident_check.txt 2>&1; echo "resident_check rc=$?"
tail -8 gpurun_out/r02l_resident_check.txt
YSM_TRACE=1 timeout 120 python scripts/latency_probe.py 360 1 > gpurun_out/r02l_trace_cfg1.txt 2>&1; echo "trace rc=$?"
tail -34 gpurun_out/r02l_trace_cfg1.txt | head -31
timeout 900 python -m pytest tests/test_gpu_resident.py tests/test_gpu_reference_slam.py tests/test_gpu_parity.py -x -q > gpurun_out/r02l_pytest.txt 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r02l_pytest.txt
