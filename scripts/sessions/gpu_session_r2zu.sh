#!/bin/bash
# r02zu (1 GPU): duplicate cells dropped inside a chunk of the stamp's candidate list (MATCH.ANY), A/B of two builds
mkdir -p gpurun_out
for rep in 1 2; do for V in dedup nodedup; do
if [ $V = nodedup ]; then export YSM_LIB=$PWD/ab/libysm_nodedup.so; else unset YSM_LIB; fi
timeout 600 python bench.py --steps 3 --warmup 2 --no-latency --no-extras --no-cpu > gpurun_out/r02zu_bench_${V}_$rep.json 2> gpurun_out/r02zu_bench_${V}_$rep.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02zu_bench_${V}_$rep.json').read().strip().splitlines()[-1])
print('$V rep=$rep value', round(d['value']), 'e2e', round(d['e2e']['value']), 'build_ms', round(d['roofline_build']['avg_launch_ms'],4))
PY
done; done
unset YSM_LIB
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
