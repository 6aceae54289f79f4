#!/bin/bash
# r02t (1 GPU): k_sweep_pruned with CTA constants made once + penalty tables, k_tile_stamp_lists: parity suite,
# in-stream kernel trace, short bench, ncu --set full of both kernels
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02t_pytest.txt 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r02t_pytest.txt
timeout 300 python scripts/kernel_trace.py > gpurun_out/r02t_trace.txt 2>&1
sed -n '/==== last call/,$p' gpurun_out/r02t_trace.txt | head -40
timeout 900 python bench.py --steps 3 --warmup 3 --no-latency --no-extras --no-cpu > gpurun_out/r02t_bench.json 2> gpurun_out/r02t_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/r02t_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02t_bench.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','timed_region_s','gpu_launches'): print(k, d.get(k))
print('e2e', d['e2e']['value'])
print('roofline', {k:d['roofline'].get(k) for k in ('achieved','frac','lsu_frac','lookups_issued_frac','share_of_step','build_share','reduce_share','avg_launch_ms')})
print('build', d['roofline_build']['frac'], d['roofline_build']['avg_launch_ms'])
PY
BENCH="python bench.py --steps 1 --warmup 1 --matches 20000 --no-latency --no-extras --no-cpu"
for K in k_sweep_pruned k_tile_stamp_lists; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 30 -c 1 -f -o gpurun_out/r02t_$K $BENCH > gpurun_out/r02t_$K.log 2>&1; echo "$K rc=$?"
done
ls -la gpurun_out/r02t*.ncu-rep
