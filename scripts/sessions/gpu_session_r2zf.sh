#!/bin/bash
# r02zf (1 GPU): native binding of Wrapper.match_scan (csrc/ysm_pyfast.c), K = 33 stamp test, equal-sized waves:
# parity suite, smoke, full default bench
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02zf_pytest.txt 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r02zf_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1200 python bench.py > gpurun_out/r02zf_bench.json 2> gpurun_out/r02zf_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/r02zf_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02zf_bench.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step'): print(k, d.get(k))
print('e2e', d['e2e']['value'])
for k in ('cfg2_sequential',): print(k, json.dumps(d.get(k, d.get(k+'_error')))[:500])
print({k:v for k,v in d.items() if 'latency' in k or 'doorbell' in k})
PY
