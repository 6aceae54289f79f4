#!/bin/bash
# r02zq (1 GPU): resident kernel, warp-aggregated tile collect: full parity suite, phase traces, latency probes
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
YSM_TRACE=1 timeout 120 python scripts/latency_probe.py 720 10 > gpurun_out/r02zq_trace_cfg2.txt 2>&1
YSM_TRACE=1 timeout 120 python scripts/latency_probe.py 360 1 > gpurun_out/r02zq_trace_cfg1.txt 2>&1
for f in cfg1 cfg2; do grep "barrier 2  \|workers stamp\|workers collect\|slowest stamp\|resident: results" gpurun_out/r02zq_trace_$f.txt | tail -5; done
for S in "360 1" "720 10" "360 1" "720 10"; do timeout 120 python scripts/latency_probe.py $S 2>&1 | grep "Wrapper.match_scan"; done
