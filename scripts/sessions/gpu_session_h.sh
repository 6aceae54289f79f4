#!/bin/bash
# r01h: throughput-path host/GPU overlap: lanes sweep + host phase trace of one step
TAG=${1:-r01h}
mkdir -p gpurun_out
nproc > gpurun_out/${TAG}_nproc.txt
for L in 2 3 4; do
  timeout 300 python bench.py --lanes $L --steps 6 --warmup 3 --no-latency --no-cpu > gpurun_out/${TAG}_bench_l$L.json 2> gpurun_out/${TAG}_bench_l$L.err; echo "bench lanes=$L rc=$?"; cat gpurun_out/${TAG}_bench_l$L.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['roofline']['avg_launch_ms'], d['roofline']['share_of_step'], d['roofline']['build_share'])"
done
YSM_TRACE=1 timeout 300 python bench.py --lanes 2 --steps 1 --warmup 3 --no-latency --no-cpu > gpurun_out/${TAG}_trace.json 2> gpurun_out/${TAG}_trace.err; echo "trace rc=$?"
tail -120 gpurun_out/${TAG}_trace.err
