#!/bin/bash
# r01f session: new register-tile stamp kernel -- parity, lanes sweep, latency phase trace, launch list
TAG=${1:-r01f}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
tail -5 gpurun_out/${TAG}_pytest.log
for L in 2 3 4; do
  timeout 300 python bench.py --lanes $L --steps 6 --warmup 3 --no-latency --no-cpu > gpurun_out/${TAG}_bench_l$L.json 2> gpurun_out/${TAG}_bench_l$L.err; echo "bench lanes=$L rc=$?"; cat gpurun_out/${TAG}_bench_l$L.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['roofline']['avg_launch_ms'], d['roofline']['share_of_step'], d['roofline']['build_share'])"
done
YSM_TRACE=1 timeout 200 python scripts/latency_probe.py 360 1 > gpurun_out/${TAG}_lat_cfg1.log 2>&1; grep -v "^\[ysm\]" gpurun_out/${TAG}_lat_cfg1.log | head -5; tail -40 gpurun_out/${TAG}_lat_cfg1.log
YSM_TRACE=1 timeout 200 python scripts/latency_probe.py 720 10 > gpurun_out/${TAG}_lat_cfg2.log 2>&1; grep -v "^\[ysm\]" gpurun_out/${TAG}_lat_cfg2.log | head -5; tail -40 gpurun_out/${TAG}_lat_cfg2.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-latency --no-cpu > gpurun_out/${TAG}_ncu_bench.log 2>&1; echo "ncu launches rc=$?"
