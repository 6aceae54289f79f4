#!/bin/bash
# r02u (1 GPU): k_sweep_pruned shared-memory budget A/B (L1 left to the lookups): in-stream kernel trace
mkdir -p gpurun_out
for KB in 96 64 48 32 24; do
  echo "--- YSM_SWEEP_SMEM_KB=$KB"
  YSM_SWEEP_SMEM_KB=$KB timeout 300 python scripts/kernel_trace.py > gpurun_out/r02u_trace_$KB.txt 2>&1
  sed -n '/==== last call/,$p' gpurun_out/r02u_trace_$KB.txt | grep "sweep_lattice\|tile_stamp\|find_valid"
done
