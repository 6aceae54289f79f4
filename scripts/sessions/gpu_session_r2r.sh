#!/bin/bash
# r02r (1 GPU): k_tile_stamp_lists (per-group / per-half step lists): parity suite, A/B kernel trace, short bench
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02r_pytest.txt 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r02r_pytest.txt
echo "--- trace new"; timeout 300 python scripts/kernel_trace.py 2>&1 | grep -i "k_\|total\|us" | head -40 | tee gpurun_out/r02r_trace_new.txt
echo "--- trace old stamp"; YSM_TRACE_DEBUG=64 timeout 300 python scripts/kernel_trace.py 2>&1 | grep -i "k_\|total\|us" | head -40 | tee gpurun_out/r02r_trace_old.txt
timeout 900 python bench.py --steps 3 --warmup 3 --no-latency --no-extras --no-cpu > gpurun_out/r02r_bench.json 2> gpurun_out/r02r_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/r02r_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02r_bench.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','timed_region_s','gpu_launches'): print(k, d.get(k))
print('e2e', d['e2e'])
print('roofline', {k:d['roofline'].get(k) for k in ('achieved','frac','lsu_frac','lookups_issued_frac','share_of_step','build_share','reduce_share','avg_launch_ms')})
print('build', d['roofline_build'])
PY
