#!/bin/bash
# r01l (1 GPU): chain-finder parity + probe
TAG=${1:-r01l}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_chains.py -m gpu -x -q > gpurun_out/${TAG}_pytest_chains.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/${TAG}_pytest_chains.log
timeout 300 python scripts/chains_probe.py > gpurun_out/${TAG}_chains_probe.json 2> gpurun_out/${TAG}_chains_probe.err; echo "probe rc=$?"; cat gpurun_out/${TAG}_chains_probe.json; tail -3 gpurun_out/${TAG}_chains_probe.err
