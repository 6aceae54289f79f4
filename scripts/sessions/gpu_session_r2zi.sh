#!/bin/bash
# r02zi (1 GPU): full default bench with the new sections (final map + ray-walk, chain finder, map matcher)
mkdir -p gpurun_out
timeout 1500 python bench.py > gpurun_out/r02zi_bench.json 2> gpurun_out/r02zi_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/r02zi_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02zi_bench.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step'): print(k, d.get(k))
print('e2e', d['e2e']['value'])
for k in ('cfg5_final_map','f2_chain_finder','f3_map_match'): print(k, json.dumps(d.get(k, d.get(k+'_error')))[:1200])
print({k:v for k,v in d.items() if 'p50_latency' in k})
PY
