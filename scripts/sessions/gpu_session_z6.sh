#!/bin/bash
# r01z6 (1 GPU): ncu --set full of ONE launch of the latency kernel (k_match_small), cfg-1 shape
mkdir -p gpurun_out
timeout 85 ncu --set full --clock-control none --import-source on -k regex:k_match_small -s 40 -c 1 -f -o gpurun_out/r01z6_k_match_small python scripts/latency_probe.py 360 1 > gpurun_out/r01z6_ncu.log 2>&1; echo "ncu rc=$?"
tail -4 gpurun_out/r01z6_ncu.log; ls -la gpurun_out/*.ncu-rep
