#!/bin/bash
# r01x (1 GPU): compute-sanitizer over the GPU suite: memcheck (everything), racecheck (grid build with exact
# candidate lists, chain finder)
TAG=${1:-r01x}
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_parity.py::test_fullsize_loop_closure_batch_properties --deselect tests/test_gpu_parity.py::test_fullsize_sequential_log_rematch_properties > gpurun_out/${TAG}_memcheck_all.log 2>&1; echo "memcheck all rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" gpurun_out/${TAG}_memcheck_all.log | head -8
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py tests/test_gpu_chains.py -m gpu -x -q -k "candidate_lists or golden" > gpurun_out/${TAG}_racecheck.log 2>&1; echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed|hazard" gpurun_out/${TAG}_racecheck.log | head -8
