#!/bin/bash
# r01y (N GPUs): weak-scaling bench at N = $1
N=${1:-2}
mkdir -p gpurun_out
nproc > gpurun_out/r01y_host_n$N.txt
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 --no-latency --no-cpu > gpurun_out/r01y_bench_n$N.json 2> gpurun_out/r01y_bench_n$N.err; echo "rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r01y_bench_n$N.json').read().strip().splitlines()[-1])
print('N', d['n_gpus'], 'value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'lanes', d['config']['lanes'], 'cores', open('gpurun_out/r01y_host_n$N.txt').read().strip())
PY
