#!/bin/bash
# r02zb (1 GPU): k_tile_clear writing whole aligned sectors: parity suite, kernel trace, bench x2
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02zb_pytest.txt 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r02zb_pytest.txt
timeout 300 python scripts/kernel_trace.py > gpurun_out/r02zb_trace.txt 2>&1
sed -n '/==== last call/,$p' gpurun_out/r02zb_trace.txt | head -14
for rep in 1 2; do
timeout 600 python bench.py --steps 3 --warmup 2 --no-latency --no-extras --no-cpu > gpurun_out/r02zb_bench_$rep.json 2> gpurun_out/r02zb_bench_$rep.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02zb_bench_$rep.json').read().strip().splitlines()[-1])
print('rep=$rep value', round(d['value']), 'e2e', round(d['e2e']['value']), 'sweep_ms', round(d['roofline']['avg_launch_ms'],4), 'build_ms', round(d['roofline_build']['avg_launch_ms'],4))
PY
done
