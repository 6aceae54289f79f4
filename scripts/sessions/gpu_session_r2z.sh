#!/bin/bash
# r02z (1 GPU): ncu launch list of the bench command (reduced) + --set full of the hot kernels at HEAD
mkdir -p gpurun_out
BENCH="python bench.py --steps 1 --warmup 1 --matches 20000 --no-latency --no-extras --no-cpu"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02z_launches.csv $BENCH > gpurun_out/r02z_bench_under_ncu.log 2>&1; echo "launch list rc=$?"
for K in k_sweep_pruned k_tile_stamp_lists k_find_valid k_tile_clear k_sweep_fine9 k_reduce; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 30 -c 1 -f -o gpurun_out/r02z_$K $BENCH > gpurun_out/r02z_$K.log 2>&1; echo "$K rc=$?"
done
ls -la gpurun_out/r02z*.ncu-rep | wc -l
