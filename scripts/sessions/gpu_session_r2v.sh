#!/bin/bash
# r02v (1 GPU): k_sweep_pruned rows per CTA A/B: in-stream kernel trace
mkdir -p gpurun_out
for R in 28 13 9 7; do
  echo "--- YSM_SWEEP_ROWS=$R"
  YSM_SWEEP_ROWS=$R timeout 300 python scripts/kernel_trace.py > gpurun_out/r02v_trace_$R.txt 2>&1
  sed -n '/==== last call/,$p' gpurun_out/r02v_trace_$R.txt | grep "sweep_lattice"
done
