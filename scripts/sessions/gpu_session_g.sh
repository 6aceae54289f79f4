#!/bin/bash
# r01g session: parity, smoke, bench (both arms), latency phase trace, launch list, ncu --set full of stamp + sweep
TAG=${1:-r01g}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
tail -5 gpurun_out/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${TAG}_smoke.log
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; cat gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err; cat gpurun_out/${TAG}_bench_reference.json
YSM_TRACE=1 timeout 200 python scripts/latency_probe.py 360 1 > gpurun_out/${TAG}_lat_cfg1.log 2>&1; grep -v "^\[ysm\]" gpurun_out/${TAG}_lat_cfg1.log | head -5; tail -40 gpurun_out/${TAG}_lat_cfg1.log
YSM_TRACE=1 timeout 200 python scripts/latency_probe.py 720 10 > gpurun_out/${TAG}_lat_cfg2.log 2>&1; grep -v "^\[ysm\]" gpurun_out/${TAG}_lat_cfg2.log | head -5; tail -40 gpurun_out/${TAG}_lat_cfg2.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-latency --no-cpu > gpurun_out/${TAG}_ncu_bench.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tile_stamp -s 4 -c 2 -o gpurun_out/${TAG}_stamp -f python bench.py --steps 2 --warmup 3 --no-latency --no-cpu > gpurun_out/${TAG}_ncu_full2.log 2>&1; echo "ncu full stamp rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sweep_pruned -s 4 -c 2 -o gpurun_out/${TAG}_sweep -f python bench.py --steps 2 --warmup 3 --no-latency --no-cpu > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "ncu full sweep rc=$?"
ls -la gpurun_out
