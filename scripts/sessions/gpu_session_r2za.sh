#!/bin/bash
# r02za (1 GPU): wave size A/B (multiple of the CTAs k_find_valid keeps resident)
mkdir -p gpurun_out
for rep in 1 2; do for W in 0 296 148; do
YSM_WAVE_ALIGN=$W timeout 600 python bench.py --steps 3 --warmup 2 --no-latency --no-extras --no-cpu > gpurun_out/r02za_bench_w${W}_$rep.json 2> gpurun_out/r02za_bench_w${W}_$rep.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02za_bench_w${W}_$rep.json').read().strip().splitlines()[-1])
print('align=$W rep=$rep value', round(d['value']), 'e2e', round(d['e2e']['value']), 'sweep_ms', round(d['roofline']['avg_launch_ms'],4), 'build_ms', round(d['roofline_build']['avg_launch_ms'],4))
PY
done; done
