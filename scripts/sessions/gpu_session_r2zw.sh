#!/bin/bash
# r02zw (1 GPU): round-2 HEAD: parity suite, smoke, ncu launch list of the (reduced) bench command + --set full of
# the hot kernels, then both bench arms un-profiled
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02zw_pytest.txt 2>&1; echo "pytest rc=$?"
tail -2 gpurun_out/r02zw_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
BENCH="python bench.py --steps 1 --warmup 1 --matches 20000 --no-latency --no-extras --no-cpu"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3200 --csv --log-file gpurun_out/r02zw_launches.csv $BENCH > gpurun_out/r02zw_bench_under_ncu.log 2>&1; echo "launch list rc=$?"
for K in k_sweep_pruned k_tile_stamp_lists k_find_valid; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 30 -c 1 -f -o gpurun_out/r02zw_$K $BENCH > gpurun_out/r02zw_$K.log 2>&1; echo "$K rc=$?"
done
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02zw_bench_reference.json 2> gpurun_out/r02zw_ref.err; echo "ref rc=$?"
timeout 1500 python bench.py > gpurun_out/r02zw_bench.json 2> gpurun_out/r02zw_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02zw_bench.json').read().strip().splitlines()[-1])
r=json.loads(open('gpurun_out/r02zw_bench_reference.json').read().strip().splitlines()[-1])
print('value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ref', round(r['value']), 'ratio', round(d['e2e']['value']/r['value'],1))
print('roofline', {k:d['roofline'][k] for k in ('frac','lsu_frac','share_of_timed_kernels','avg_launch_ms')}, 'build', round(d['roofline_build']['frac'],3), 'step', round(d['roofline_step']['frac'],3))
print({k:v for k,v in d.items() if 'p50_latency' in k or 'speedup' in k})
for k in ('cfg3','cfg4','cfg2_sequential','cfg5_final_map','f2_chain_finder','f3_map_match'): print(k, 'error' if k+'_error' in d else 'ok')
PY
