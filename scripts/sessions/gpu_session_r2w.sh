#!/bin/bash
# r02w (1 GPU): k_sweep_fine9, stamp write-out / list capacity, one-round sweep prologue: parity suite, kernel
# trace, short bench at 3 and 4 lanes
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02w_pytest.txt 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r02w_pytest.txt
timeout 300 python scripts/kernel_trace.py > gpurun_out/r02w_trace.txt 2>&1
sed -n '/==== last call/,$p' gpurun_out/r02w_trace.txt | head -40
for L in 3 4; do
timeout 900 python bench.py --steps 3 --warmup 3 --lanes $L --no-latency --no-extras --no-cpu > gpurun_out/r02w_bench_l$L.json 2> gpurun_out/r02w_bench_l$L.err; echo "bench lanes=$L rc=$?"
tail -3 gpurun_out/r02w_bench_l$L.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02w_bench_l$L.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step'): print(k, d.get(k))
print('e2e', d['e2e']['value'])
print('roofline', {k:d['roofline'].get(k) for k in ('achieved','frac','lsu_frac','share_of_step','build_share','reduce_share','avg_launch_ms')})
print('build', d['roofline_build']['frac'], d['roofline_build']['avg_launch_ms'])
PY
done
