#!/bin/bash
# r02zs (1 GPU): HEAD: parity suite, smoke, both bench arms (all sections)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02zs_pytest.txt 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r02zs_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02zs_bench_reference.json 2> gpurun_out/r02zs_ref.err; echo "ref rc=$?"
cut -c1-200 gpurun_out/r02zs_bench_reference.json
timeout 1500 python bench.py > gpurun_out/r02zs_bench.json 2> gpurun_out/r02zs_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/r02zs_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02zs_bench.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','timed_region_s','gpu_launches'): print(k, d.get(k))
print('e2e', d['e2e']['value'])
print('roofline', {k:d['roofline'][k] for k in ('achieved','frac','lsu_frac','share_of_timed_kernels','share_of_step','build_share','avg_launch_ms')})
print('build', d['roofline_build']['frac'], 'step', d['roofline_step']['frac'], 'cpu', d['cpu_baseline']['value'])
for k in ('cfg2_rematch','cfg3','cfg4','cfg2_sequential','cfg5_final_map','f2_chain_finder','f3_map_match'): print(k, json.dumps(d.get(k, d.get(k+'_error')))[:330])
print({k:v for k,v in d.items() if 'latency' in k or 'doorbell' in k})
PY
