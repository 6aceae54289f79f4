#!/bin/bash
# r02zz (8 GPUs): N = 8 and N = 1 at round-2 HEAD (strong scaling of the 100k relocalisation batch)
mkdir -p gpurun_out
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r02zz_bench_n8.json 2> gpurun_out/r02zz_bench_n8.err; echo "N=8 rc=$?"
timeout 100 python bench.py --steps 5 --warmup 3 --no-latency --no-extras --no-cpu > gpurun_out/r02zz_bench_n1.json 2> gpurun_out/r02zz_bench_n1.err; echo "N=1 rc=$?"
python - <<'PY'
import json
for n in (8,1):
    d=json.loads(open('gpurun_out/r02zz_bench_n%d.json'%n).read().strip().splitlines()[-1])
    print('N', d['n_gpus'], 'value', round(d['value']), 'ms/step', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']))
PY
