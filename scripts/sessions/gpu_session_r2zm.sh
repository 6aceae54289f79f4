#!/bin/bash
# r02zm (8 GPUs): strong scaling of the headline workload at HEAD, equal-sized waves (100k relocalisation queries sharded over the ranks)
mkdir -p gpurun_out
nproc
for N in 8 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r02zm_bench_n$N.json 2> gpurun_out/r02zm_bench_n$N.err; echo "N=$N rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02zm_bench_n$N.json').read().strip().splitlines()[-1])
    print('N', d['n_gpus'], 'value', round(d['value']), 'ms/step', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), 'lanes', d['config']['lanes'], 'sweep avg ms', d['roofline']['avg_launch_ms'])
except Exception as e:
    print('parse failed', e); print(open('gpurun_out/r02zm_bench_n$N.err').read()[-1500:])
PY
done
timeout 600 python bench.py --steps 5 --warmup 3 --no-latency --no-extras --no-cpu > gpurun_out/r02zm_bench_n1.json 2> gpurun_out/r02zm_bench_n1.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02zm_bench_n1.json').read().strip().splitlines()[-1])
print('N', d['n_gpus'], 'value', round(d['value']), 'ms/step', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']))
PY
