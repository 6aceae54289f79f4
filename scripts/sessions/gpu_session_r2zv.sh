#!/bin/bash
# r02zv (1 GPU): 3 vs 4 lanes with the same wave size (344 matches): 16.5 GiB / 3 lanes vs 22 GiB / 4 lanes
mkdir -p gpurun_out
for rep in 1 2; do for V in "3 16.5" "4 22" "4 30"; do
set -- $V
timeout 600 python bench.py --steps 3 --warmup 2 --lanes $1 --grid-gb $2 --no-latency --no-extras --no-cpu > gpurun_out/r02zv_bench_l$1_g$2_$rep.json 2> gpurun_out/r02zv_bench_l$1_g$2_$rep.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02zv_bench_l$1_g$2_$rep.json').read().strip().splitlines()[-1])
print('lanes=$1 GiB=$2 rep=$rep value', round(d['value']), 'e2e', round(d['e2e']['value']), 'launches', d['roofline']['launches'])
PY
done; done
