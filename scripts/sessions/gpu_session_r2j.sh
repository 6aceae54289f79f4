#!/bin/bash
# r02j (1 GPU): warp-wide doorbell poll, all-ties fast path in the reduce (cfg 3), full test-suite + bench
# reference arm
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02j_pytest.txt 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r02j_pytest.txt
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02j_bench.json 2> gpurun_out/r02j_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/r02j_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02j_bench.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','timed_region_s','workload_generate_s','gpu_launches'): print(k, d.get(k))
print('e2e', d['e2e'])
print('roofline', {k:d['roofline'][k] for k in ('achieved','frac','lsu_frac','lookups_issued_frac','share_of_step','build_share','reduce_share','avg_launch_ms')})
print('build', d['roofline_build']); print('step', d['roofline_step']); print('cpu', d['cpu_baseline'])
for k in ('cfg2_rematch','cfg3','cfg4','cfg2_sequential'): print(k, json.dumps(d.get(k, d.get(k+'_error')))[:900])
print({k:v for k,v in d.items() if 'latency' in k or 'doorbell' in k})
PY
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02j_bench_reference.json 2> gpurun_out/r02j_ref.err; echo "ref rc=$?"
cut -c1-400 gpurun_out/r02j_bench_reference.json
