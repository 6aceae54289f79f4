#!/bin/bash
# r02y (1 GPU): same-box A/B of the headline bench: new kernels vs YSM_DEBUG_OR=192 (k_tile_stamp + k_sweep_points)
mkdir -p gpurun_out
for rep in 1 2; do for F in 0 192 64 128; do
YSM_DEBUG_OR=$F timeout 600 python bench.py --steps 3 --warmup 2 --no-latency --no-extras --no-cpu > gpurun_out/r02y_bench_f${F}_$rep.json 2> gpurun_out/r02y_bench_f${F}_$rep.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02y_bench_f${F}_$rep.json').read().strip().splitlines()[-1])
print('flags=$F rep=$rep value', round(d['value']), 'e2e', round(d['e2e']['value']), 'sweep_ms', round(d['roofline']['avg_launch_ms'],4), 'build_ms', round(d['roofline_build']['avg_launch_ms'],4))
PY
done; done
