#!/bin/bash
# r02zy (1 GPU): fine reduce takes the angular-covariance sums from k_sweep_fine9's integer sums (no re-gather):
# parity suite, in-stream kernel trace, bench x2
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 300 python scripts/kernel_trace.py > gpurun_out/r02zy_trace.txt 2>&1
sed -n '/==== last call/,$p' gpurun_out/r02zy_trace.txt | grep "k_reduce\|k_sweep_points"
for rep in 1 2; do
timeout 600 python bench.py --steps 3 --warmup 2 --no-latency --no-extras --no-cpu > gpurun_out/r02zy_bench_$rep.json 2> gpurun_out/r02zy_bench_$rep.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02zy_bench_$rep.json').read().strip().splitlines()[-1])
print('rep=$rep value', round(d['value']), 'e2e', round(d['e2e']['value']))
PY
done
