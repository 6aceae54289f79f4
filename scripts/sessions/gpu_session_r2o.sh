#!/bin/bash
# r02o (1 GPU): full GPU test-suite after the degenerate-schedule / raytrace-workspace changes
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02o_pytest.txt 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/r02o_pytest.txt
