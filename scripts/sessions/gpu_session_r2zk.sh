#!/bin/bash
# r02zk (1 GPU): wave size A/B (slots per rank): does an L2-resident working set pay for the smaller waves?
mkdir -p gpurun_out
for S in 0 768 512 384 256; do
timeout 600 python bench.py --steps 3 --warmup 2 --max-slots $S --no-latency --no-extras --no-cpu > gpurun_out/r02zk_bench_s$S.json 2> gpurun_out/r02zk_bench_s$S.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02zk_bench_s$S.json').read().strip().splitlines()[-1])
print('slots=$S value', round(d['value']), 'e2e', round(d['e2e']['value']), 'sweep_ms', round(d['roofline']['avg_launch_ms'],4), 'build_ms', round(d['roofline_build']['avg_launch_ms'],4), 'launches', d['roofline']['launches'])
PY
done
