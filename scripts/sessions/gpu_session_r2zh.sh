#!/bin/bash
# r02zh (1 GPU): compute-sanitizer at round-2 HEAD. The resident kernel spins on a doorbell and is not run under
# the tool (YSM_NO_RESIDENT=1: single queries take k_match_small); the tests that assert it ran are deselected.
mkdir -p gpurun_out
export YSM_NO_RESIDENT=1
timeout 1300 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests -m gpu -q --deselect tests/test_gpu_parity.py::test_fullsize_loop_closure_batch_properties --deselect tests/test_gpu_parity.py::test_fullsize_sequential_log_rematch_properties --deselect tests/test_gpu_parity.py::test_latency_path_device_chained_fine_pass --ignore tests/test_gpu_resident.py --ignore tests/test_gpu_relocalisation.py --ignore tests/test_gpu_reference_slam.py > gpurun_out/r02zh_memcheck_all.log 2>&1; echo "memcheck all rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" gpurun_out/r02zh_memcheck_all.log | head -8
