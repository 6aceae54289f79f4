#!/bin/bash
# r02zd (1 GPU): HEAD: parity suite, smoke, full default bench (both arms)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02zd_pytest.txt 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r02zd_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02zd_bench_reference.json 2> gpurun_out/r02zd_ref.err; echo "ref rc=$?"
cut -c1-300 gpurun_out/r02zd_bench_reference.json
timeout 1200 python bench.py > gpurun_out/r02zd_bench.json 2> gpurun_out/r02zd_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/r02zd_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02zd_bench.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','timed_region_s','workload_generate_s','gpu_launches'): print(k, d.get(k))
print('e2e', d['e2e'])
print('roofline', {k:d['roofline'][k] for k in ('achieved','frac','lsu_frac','lookups_issued_frac','share_of_step','build_share','reduce_share','avg_launch_ms')})
print('build', d['roofline_build']['frac'], 'step', d['roofline_step']['frac'], 'cpu', d['cpu_baseline'])
for k in ('cfg2_rematch','cfg3','cfg4','cfg2_sequential'): print(k, json.dumps(d.get(k, d.get(k+'_error')))[:700])
print({k:v for k,v in d.items() if 'latency' in k or 'doorbell' in k})
PY
