#!/bin/bash
# r01z3 (1 GPU): verify the single-query glue change -- GPU suite, latency probes with phase trace, bench
TAG=${1:-r01z3}
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest.log
(timeout 60 python scripts/latency_probe.py 360 1; YSM_TRACE=1 timeout 60 python scripts/latency_probe.py 360 1 2>&1 | tail -28; timeout 60 python scripts/latency_probe.py 720 10) > gpurun_out/${TAG}_latency.txt 2>&1; echo "probe rc=$?"; grep "p50" gpurun_out/${TAG}_latency.txt
timeout 300 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in d if k.startswith(('value','p50','p99','latency_'))}, d['e2e']['value'])
PY
tail -3 gpurun_out/${TAG}_bench.err
