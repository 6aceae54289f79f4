#!/bin/bash
# r02n (1 GPU): A/B of the resident kernel's CTA size: 1024 threads (default build) vs 512 (libysm_b200_t512.so)
mkdir -p gpurun_out
for L in "" yag-slam_b200/csrc/libysm_b200_t512.so; do
  echo "== YSM_LIB=$L"
  YSM_LIB=$L timeout 200 python scripts/resident_check.py 2>&1 | grep -E "Wrapper|cold|ALL OK|FAIL|bad=[1-9]"
done
