#!/bin/bash
# r01z7 (1 GPU): lanes A/B on one box (3 vs 4), throughput arms only
mkdir -p gpurun_out
for L in 3 4; do
  timeout 40 python bench.py --lanes $L --steps 10 --warmup 3 --no-latency --no-cpu > gpurun_out/r01z7_bench_l$L.json 2> gpurun_out/r01z7_bench_l$L.err; echo "lanes $L rc=$?"
  python - <<PY
import json
d=json.loads(open('gpurun_out/r01z7_bench_l$L.json').read().strip().splitlines()[-1])
print('lanes', d['config']['lanes'], 'value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']))
PY
done
