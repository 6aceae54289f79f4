#!/bin/bash
# r02zn (8 GPUs): N = 8 twice more (the r02zm e2e figure at N = 8 was an outlier?)
mkdir -p gpurun_out
for rep in 1 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r02zn_bench_n8_$rep.json 2> gpurun_out/r02zn_bench_n8_$rep.err; echo "rc=$?"
  python - <<PY
import json
d=json.loads(open('gpurun_out/r02zn_bench_n8_$rep.json').read().strip().splitlines()[-1])
print('N', d['n_gpus'], 'value', round(d['value']), 'ms/step', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), 'e2e ms', round(d['e2e']['ms_per_step'],2))
PY
done
