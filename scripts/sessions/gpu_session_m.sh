#!/bin/bash
# r01m (1 GPU): map-matching + checkpoint re-matching parity, probes
TAG=${1:-r01m}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_map.py tests/test_gpu_graph_io.py -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/${TAG}_pytest.log
timeout 300 python scripts/map_probe.py > gpurun_out/${TAG}_map_probe.json 2> gpurun_out/${TAG}_map_probe.err; echo "probe rc=$?"; cat gpurun_out/${TAG}_map_probe.json; tail -3 gpurun_out/${TAG}_map_probe.err
