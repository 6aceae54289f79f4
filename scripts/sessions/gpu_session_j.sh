#!/bin/bash
# r01j (1 GPU): parity at HEAD, full default bench + reference arm, latency phase trace (cfg1/cfg2), launch list
TAG=${1:-r01j}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${TAG}_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest.log
timeout 500 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; cat gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; echo "ref rc=$?"; cat gpurun_out/${TAG}_bench_reference.json
for cfg in "360 1" "720 10"; do
  echo "== latency probe $cfg" >> gpurun_out/${TAG}_latency_trace.txt
  YSM_TRACE=1 timeout 200 python scripts/latency_probe.py $cfg >> gpurun_out/${TAG}_latency_trace.txt 2>&1
done
tail -60 gpurun_out/${TAG}_latency_trace.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-latency --no-cpu > gpurun_out/${TAG}_ncu_bench.log 2>&1; echo "ncu launches rc=$?"
