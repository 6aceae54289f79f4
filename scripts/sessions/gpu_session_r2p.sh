#!/bin/bash
# r02p (1 GPU): ncu launch list of the bench command (reduced step count) + --set full of the three hot kernels
mkdir -p gpurun_out
BENCH="python bench.py --steps 1 --warmup 1 --matches 20000 --no-latency --no-extras --no-cpu"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02p_launches.csv $BENCH > gpurun_out/r02p_bench_under_ncu.log 2>&1; echo "launch list rc=$?"
for K in k_sweep_pruned k_tile_stamp k_find_valid k_reduce k_sweep_points; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 30 -c 1 -f -o gpurun_out/r02p_$K $BENCH > gpurun_out/r02p_$K.log 2>&1; echo "$K rc=$?"
done
ls -la gpurun_out/*.ncu-rep | tail -6
