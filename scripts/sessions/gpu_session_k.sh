#!/bin/bash
# r01k (1 GPU): ncu --set full with source for the three grid-build / sweep kernels of the bench workload
TAG=${1:-r01k}
mkdir -p gpurun_out
for k in k_tile_stamp k_sweep_pruned k_find_valid; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -o gpurun_out/${TAG}_$k -f python bench.py --steps 2 --warmup 3 --no-latency --no-cpu > gpurun_out/${TAG}_ncu_$k.log 2>&1; echo "ncu $k rc=$?"
done
ls -la gpurun_out
